/*
 * grok_b200/csrc/host_pack.cpp -- host-side sample-container conversion for the PCIe legs.
 *
 * Grok hands the plugin 32-bit sample planes (grk_image_comp::data, grok.h; gpup_image_comp in
 * plugin/gpup/gpu_plugin_shared.h L34-48) although JPEG 2000 samples of up to 16 bits fit half
 * of that.  PCIe is the end-to-end bound of b2k_encode / b2k_decode, so the int32
 * entry points narrow each pipeline chunk into a pinned 16-bit staging buffer on a small pool of
 * host threads while the previous chunk crosses the bus, and widen on the way back.  A host that
 * already owns pinned 16-bit planes calls b2k_encode16 / b2k_decode16 and skips this file.
 *
 * Pure host code: no CUDA, no reference code.  Nothing here touches coefficient values -- it is
 * a container change (truncate to 16 bits / zero- or sign-extend), so parity is unaffected.
 */
#include "b2k_internal.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sched.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace
{

/* fork-join pool: parallel_for(n, fn) runs fn(i) for i in [0, n) on the workers plus the caller.
   Inside a session (begin_session / end_session: one b2k_encode / b2k_decode call) idle workers spin on the
   generation counter instead of sleeping, so a fork-join costs a few microseconds -- the calls issue ~100 of them. */
class HostPool
{
public:
  explicit HostPool(int nthreads)
  {
    const std::vector<int> cpus = spread_cpus();
    for(int i = 0; i < nthreads - 1; ++i)
    {
      workers_.emplace_back([this] { loop(); });
      if(!cpus.empty())
      { /* one worker per physical core first: two bandwidth-bound threads on SMT siblings gain nothing */
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(cpus[(size_t)(i + 1) % cpus.size()], &set);
        pthread_setaffinity_np(workers_.back().native_handle(), sizeof(set), &set);
      }
    }
  }
  ~HostPool()
  {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_.store(true);
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for(auto& t : workers_)
      t.join();
  }
  int size() const { return (int)workers_.size() + 1; }
  void begin_session() { hot_.fetch_add(1, std::memory_order_relaxed); }
  void end_session() { hot_.fetch_sub(1, std::memory_order_relaxed); }
  void parallel_for(size_t n, const std::function<void(size_t)>& fn)
  {
    if(n == 0)
      return;
    if(workers_.empty() || n == 1)
    {
      for(size_t i = 0; i < n; ++i)
        fn(i);
      return;
    }
    /* one fork-join at a time: engines on several GPUs (or a codestream writer) may call from different threads */
    std::lock_guard<std::mutex> call(call_mu_);
    fn_ = &fn;
    n_ = n;
    next_.store(0, std::memory_order_relaxed);
    pending_.store((int)workers_.size(), std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_); /* pairs with the sleepers' predicate check */
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    drain();
    /* every worker must have left this generation before fn goes out of scope; they are at most one task away */
    for(int spins = 0; pending_.load(std::memory_order_acquire) != 0;)
    {
      cpu_relax();
      if((++spins & 127) == 0)
        sched_yield();
    }
    fn_ = nullptr;
  }

private:
  /* how long an idle worker polls for the next fork-join of a session before it sleeps (B2K_HOST_SPIN_US) */
  static int spin_limit()
  {
    static const int v = [] {
      const char* e = getenv("B2K_HOST_SPIN_US");
      const int us = e ? std::max(0, atoi(e)) : 100;
      return us * 25; /* ~40 ns per pause */
    }();
    return v;
  }
  static void cpu_relax()
  {
#if defined(__x86_64__)
    _mm_pause();
#else
    std::this_thread::yield();
#endif
  }
  /* CPUs this process may use, ordered so that distinct physical cores come first (B2K_HOST_PIN=0: no pinning) */
  static std::vector<int> spread_cpus()
  {
    std::vector<int> out;
    const char* pin = getenv("B2K_HOST_PIN");
    if(pin && atoi(pin) == 0)
      return out;
    cpu_set_t set;
    if(sched_getaffinity(0, sizeof(set), &set) != 0)
      return out;
    std::vector<std::pair<long, int>> first, rest; /* (package << 16 | core, cpu) */
    std::vector<long> seen;
    for(int cpu = 0; cpu < CPU_SETSIZE; ++cpu)
    {
      if(!CPU_ISSET(cpu, &set))
        continue;
      long core = cpu, pkg = 0;
      char path[128];
      snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/core_id", cpu);
      if(FILE* f = fopen(path, "r"))
      {
        if(fscanf(f, "%ld", &core) != 1) core = cpu;
        fclose(f);
      }
      snprintf(path, sizeof(path), "/sys/devices/system/cpu/cpu%d/topology/physical_package_id", cpu);
      if(FILE* f = fopen(path, "r"))
      {
        if(fscanf(f, "%ld", &pkg) != 1) pkg = 0;
        fclose(f);
      }
      const long key = (pkg << 16) | core;
      bool dup = false;
      for(long k : seen)
        dup |= (k == key);
      if(dup)
        rest.push_back({key, cpu});
      else
      {
        seen.push_back(key);
        first.push_back({key, cpu});
      }
    }
    for(auto& p : first) out.push_back(p.second);
    for(auto& p : rest) out.push_back(p.second);
    return out;
  }
  void drain()
  {
    for(;;)
    {
      const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
      if(i >= n_)
        break;
      (*fn_)(i);
    }
  }
  void loop()
  {
    uint64_t seen = 0;
    for(;;)
    {
      uint64_t g;
      int spins = 0;
      while((g = gen_.load(std::memory_order_acquire)) == seen)
      {
        if(hot_.load(std::memory_order_relaxed) > 0 && spins < spin_limit())
        {
          cpu_relax();
          if((++spins & 127) == 0)
            sched_yield(); /* a spinner must never keep the caller's (or the CUDA runtime's) threads off its core */
          continue;
        }
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait_for(lk, std::chrono::milliseconds(hot_.load() > 0 ? 1 : 1000),
                     [&] { return gen_.load(std::memory_order_acquire) != seen; });
        spins = 0;
      }
      if(stop_.load())
        return;
      seen = g;
      drain();
      pending_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::mutex call_mu_;
  std::condition_variable cv_;
  std::atomic<bool> stop_{false};
  std::atomic<int> hot_{0};
  std::atomic<uint64_t> gen_{0};
  const std::function<void(size_t)>* fn_ = nullptr;
  std::atomic<size_t> next_{0};
  size_t n_ = 0;
  std::atomic<int> pending_{0};
};

std::mutex g_pool_mu;
HostPool* g_pool = nullptr;
int g_threads = -1; /* -1: not decided yet; 0: host packing disabled */

/* CPUs' worth of time this process may use per period, <= 0 if unlimited / unknown (cgroup v2, then v1) */
double cgroup_cpu_quota()
{
  if(FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r"))
  {
    char q[64] = {0};
    long period = 0;
    const int n = fscanf(f, "%63s %ld", q, &period);
    fclose(f);
    if(n == 2 && period > 0 && strcmp(q, "max") != 0)
      return atof(q) / (double)period;
    if(n >= 1)
      return 0;
  }
  long quota = -1, period = 0;
  if(FILE* f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r"))
  {
    if(fscanf(f, "%ld", &quota) != 1) quota = -1;
    fclose(f);
  }
  if(FILE* f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r"))
  {
    if(fscanf(f, "%ld", &period) != 1) period = 0;
    fclose(f);
  }
  return (quota > 0 && period > 0) ? (double)quota / (double)period : 0;
}

int default_threads()
{
  if(const char* s = getenv("B2K_HOST_THREADS"))
    return std::max(0, atoi(s));
  cpu_set_t set;
  int avail = (int)std::thread::hardware_concurrency();
  if(sched_getaffinity(0, sizeof(set), &set) == 0)
    avail = CPU_COUNT(&set);
  /* a CPU-time quota (cgroup cpu.max, i.e. a container's --cpus) counts before the CPU list does: a pool that
     burns more than the quota gets the whole process throttled for the rest of the 100 ms period.  Keep three
     CPUs' worth for the caller, the CUDA runtime's threads and whatever else lives in the container. */
  double quota = cgroup_cpu_quota();
  const int peers = b2k_host_local_peers();
  if(quota > 0)
    avail = std::min(avail, std::max(1, (int)(quota / peers) - 3));
  else if(peers > 1)
    avail = std::max(1, avail / peers - 1);
  /* a container change is bandwidth work: a couple of dozen cores saturate one socket's DRAM */
  return std::max(1, std::min(avail, 24));
}

HostPool* pool()
{
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if(g_threads < 0)
    g_threads = default_threads();
  if(g_threads == 0)
    return nullptr;
  if(!g_pool || g_pool->size() != g_threads)
  {
    delete g_pool;
    g_pool = new HostPool(g_threads);
  }
  return g_pool;
}

/* ---- row kernels -------------------------------------------------------------------------------- */
void narrow_row_scalar(const int32_t* s, uint16_t* d, size_t n)
{
  for(size_t i = 0; i < n; ++i)
    d[i] = (uint16_t)s[i];
}
void widen_row_scalar(const uint16_t* s, int32_t* d, size_t n, bool sgnd)
{
  if(sgnd)
    for(size_t i = 0; i < n; ++i)
      d[i] = (int16_t)s[i];
  else
    for(size_t i = 0; i < n; ++i)
      d[i] = s[i];
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void narrow_row_avx2_cached(const int32_t* s, uint16_t* d, size_t n)
{ /* plain stores: the destination is about to be read by the device while still in cache */
  size_t i = 0;
  const __m256i m = _mm256_set1_epi32(0xFFFF);
  for(; i + 16 <= n; i += 16)
  {
    const __m256i a = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(s + i)), m);
    const __m256i b = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(s + i + 8)), m);
    _mm256_storeu_si256((__m256i*)(d + i), _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8));
  }
  for(; i < n; ++i)
    d[i] = (uint16_t)s[i];
}
__attribute__((target("avx2"))) void narrow_row_avx2(const int32_t* s, uint16_t* d, size_t n)
{
  size_t i = 0;
  while(i < n && ((uintptr_t)(d + i) & 31)) /* align the streaming stores */
  {
    d[i] = (uint16_t)s[i];
    ++i;
  }
  const __m256i m = _mm256_set1_epi32(0xFFFF);
  for(; i + 16 <= n; i += 16)
  {
    const __m256i a = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(s + i)), m);
    const __m256i b = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(s + i + 8)), m);
    const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);
    _mm256_stream_si256((__m256i*)(d + i), p);
  }
  for(; i < n; ++i)
    d[i] = (uint16_t)s[i];
}
__attribute__((target("avx2"))) void widen_row_avx2(const uint16_t* s, int32_t* d, size_t n, bool sgnd)
{
  size_t i = 0;
  while(i < n && ((uintptr_t)(d + i) & 31))
  {
    d[i] = sgnd ? (int32_t)(int16_t)s[i] : (int32_t)s[i];
    ++i;
  }
  if(sgnd)
    for(; i + 8 <= n; i += 8)
      _mm256_stream_si256((__m256i*)(d + i), _mm256_cvtepi16_epi32(_mm_loadu_si128((const __m128i*)(s + i))));
  else
    for(; i + 8 <= n; i += 8)
      _mm256_stream_si256((__m256i*)(d + i), _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i*)(s + i))));
  for(; i < n; ++i)
    d[i] = sgnd ? (int32_t)(int16_t)s[i] : (int32_t)s[i];
}
bool have_avx2()
{
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}
#endif

constexpr size_t ROWS_PER_TASK = 8;

} // namespace

/* processes sharing this host with us, one per GPU, as torchrun / mpirun advertise them (1 if unknown) */
int b2k_host_local_peers(void)
{
  for(const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS"})
    if(const char* e = getenv(name))
      return std::max(1, atoi(e));
  return 1;
}

void b2k_host_set_threads(int n)
{
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_threads = n < 0 ? default_threads() : n;
}

int b2k_host_threads(void)
{
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if(g_threads < 0)
    g_threads = default_threads();
  return g_threads;
}

void b2k_host_parallel(size_t n, const std::function<void(size_t)>& fn)
{
  if(HostPool* P = pool())
    P->parallel_for(n, fn);
  else
    for(size_t i = 0; i < n; ++i)
      fn(i);
}

void b2k_host_session(bool begin)
{
  HostPool* P = pool();
  if(!P)
    return;
  if(begin)
    P->begin_session();
  else
    P->end_session();
}

void b2k_host_convert(const b2k_host_rect* rects, size_t nrects, bool widen, bool sgnd, bool cached_dst)
{
  /* one fork-join over every rectangle of the chunk: tasks are groups of ROWS_PER_TASK rows */
  std::vector<size_t> first(nrects + 1, 0);
  for(size_t r = 0; r < nrects; ++r)
    first[r + 1] = first[r] + (rects[r].h + ROWS_PER_TASK - 1) / ROWS_PER_TASK;
  auto body = [&](size_t t) {
    size_t r = 0;
    while(t >= first[r + 1])
      ++r;
    const b2k_host_rect& R = rects[r];
    const size_t y0 = (t - first[r]) * ROWS_PER_TASK, y1 = std::min(R.h, y0 + ROWS_PER_TASK);
    for(size_t y = y0; y < y1; ++y)
    {
      if(widen)
      {
        const uint16_t* s = (const uint16_t*)R.src + y * R.src_stride;
        int32_t* d = (int32_t*)R.dst + y * R.dst_stride;
#if defined(__x86_64__)
        if(have_avx2())
        {
          widen_row_avx2(s, d, R.w, sgnd);
          continue;
        }
#endif
        widen_row_scalar(s, d, R.w, sgnd);
      }
      else
      {
        const int32_t* s = (const int32_t*)R.src + y * R.src_stride;
        uint16_t* d = (uint16_t*)R.dst + y * R.dst_stride;
#if defined(__x86_64__)
        if(have_avx2())
        {
          if(cached_dst)
            narrow_row_avx2_cached(s, d, R.w);
          else
            narrow_row_avx2(s, d, R.w);
          continue;
        }
#endif
        narrow_row_scalar(s, d, R.w);
      }
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
  };
  HostPool* P = pool();
  if(P)
    P->parallel_for(first[nrects], body);
  else
    for(size_t t = 0; t < first[nrects]; ++t)
      body(t);
}
