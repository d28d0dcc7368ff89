/*
 * grok_b200/csrc/engine.cu -- the B200 tile engine behind include/grok_b200.h.
 *
 * Replaces, for the tiles it is given, the per-tile pipeline of the reference
 *   encode: TileProcessorCompress::preCompressTile / buildCompressDAG / doCompress
 *           (tile_processor/TileProcessorCompress.cpp L104-254, L347-531, L539-)  up to, not
 *           including, rate allocation and T2;
 *   decode: TileProcessor::scheduleAndRunDecompress (tile_processor/TileProcessor.cpp L1272-)
 *           after the T2 parse.
 * The Taskflow DAG (dcShift -> MCT -> per-level vert/horiz -> T1) becomes stream-ordered kernel
 * launches over ALL selected tiles at once: one launch per decomposition level, one for the
 * block coder.  Buffers are image-shaped planes in HBM addressed by canvas coordinate, so a
 * tile is a view, tiles of every size batch into the same launch, and the engine never copies
 * a tile out of the image (the reference's preCompressTile row copy disappears).
 *
 * Product code: fails loudly (negative return + b2k_last_error) without a CUDA device; there is
 * no CPU fallback here -- the host (Grok) owns that decision via the >0 return convention.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <unordered_map>
#include <string>
#include <vector>

#include "b2k_internal.h"
#include "geometry.h"

using namespace b2k;

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_pack_policy{-1}; /* host packing: -1 auto (PackTuner), 0 never, 1 always */
static std::atomic<int> g_last_pack[2]{{-1}, {-1}};
void b2k_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }

#define CUDA_TRY(expr)                                                                             \
  do                                                                                               \
  {                                                                                                \
    cudaError_t _e = (expr);                                                                       \
    if(_e != cudaSuccess)                                                                          \
    {                                                                                              \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                                  \
      return -1;                                                                                   \
    }                                                                                              \
  } while(0)

struct b2k_engine
{
  int device = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr, h2d_stream = nullptr;
  cudaStream_t aux[4] = {nullptr, nullptr, nullptr, nullptr}; /* latency-bound side kernels run concurrently here */
  cudaDeviceProp prop{};
  /* b2k_encode / b2k_decode keep the job (device buffers, plans) of the last coding they saw.  The cache and its lock
     belong to the engine: calls on one engine are serialised, engines (one per GPU) run side by side */
  std::mutex mu;
  b2k_device_job* cached = nullptr;
};

/* ---- device memory cache --------------------------------------------------------------------------------------
 * Jobs come and go with the coding (every windowed decode is a new virtual image) and cudaMalloc /
 * cudaFree synchronise the device and cost milliseconds per gigabyte.  Freed job buffers are kept per device and
 * handed out again to requests of about the same size (best fit, at most 25 % slack); the cache is trimmed, largest
 * first, above B2K_DEV_CACHE_GB (default 32). */
namespace {
struct DevCache
{
  std::mutex mu;
  std::multimap<size_t, void*> free_;
  std::unordered_map<void*, size_t> live;
  size_t cached = 0;
};
DevCache g_devcache[32];
size_t dev_cache_limit()
{
  static const size_t v = [] {
    const char* e = getenv("B2K_DEV_CACHE_GB");
    return (size_t)(e ? atof(e) : 32.0) << 30;
  }();
  return v;
}
cudaError_t dev_alloc(void** p, size_t n)
{
  int dev = 0;
  cudaGetDevice(&dev);
  DevCache& C = g_devcache[dev & 31];
  const size_t gran = std::max<size_t>(256u << 10, n >> 4);
  const size_t want = (n + gran - 1) / gran * gran;
  {
    std::lock_guard<std::mutex> lk(C.mu);
    auto it = C.free_.lower_bound(want);
    if(it != C.free_.end() && it->first <= want + want / 4)
    {
      *p = it->second;
      C.live[*p] = it->first;
      C.cached -= it->first;
      C.free_.erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(p, want);
  if(e != cudaSuccess)
  { /* give the cache back and try once more */
    (void)cudaGetLastError();
    std::vector<void*> drop;
    {
      std::lock_guard<std::mutex> lk(C.mu);
      for(auto& kv : C.free_)
        drop.push_back(kv.second);
      C.free_.clear();
      C.cached = 0;
    }
    for(void* q : drop)
      cudaFree(q);
    e = cudaMalloc(p, want);
  }
  if(e == cudaSuccess)
  {
    std::lock_guard<std::mutex> lk(C.mu);
    C.live[*p] = want;
  }
  return e;
}
cudaError_t dev_free(void* p)
{
  if(!p)
    return cudaSuccess;
  int dev = 0;
  cudaGetDevice(&dev);
  DevCache& C = g_devcache[dev & 31];
  std::vector<void*> drop;
  {
    std::lock_guard<std::mutex> lk(C.mu);
    auto it = C.live.find(p);
    if(it == C.live.end())
      drop.push_back(p); /* not ours (or another device's): plain free */
    else
    {
      C.free_.insert({it->second, p});
      C.cached += it->second;
      C.live.erase(it);
      while(C.cached > dev_cache_limit() && !C.free_.empty())
      {
        auto big = std::prev(C.free_.end());
        drop.push_back(big->second);
        C.cached -= big->first;
        C.free_.erase(big);
      }
    }
  }
  for(void* q : drop)
    cudaFree(q);
  return cudaSuccess;
}
} // namespace
template <class T>
static inline cudaError_t dev_alloc_t(T** p, size_t n) { return dev_alloc(reinterpret_cast<void**>(p), n); }

/* the same for pinned host memory (descriptor tables, staging rings): cudaHostAlloc / cudaFreeHost are slower still */
namespace {
struct HostCache
{
  std::mutex mu;
  std::multimap<size_t, void*> free_;
  std::unordered_map<void*, size_t> live;
  size_t cached = 0;
} g_hostcache;
cudaError_t host_alloc(void** p, size_t n)
{
  const size_t gran = std::max<size_t>(64u << 10, n >> 4);
  const size_t want = (n + gran - 1) / gran * gran;
  {
    std::lock_guard<std::mutex> lk(g_hostcache.mu);
    auto it = g_hostcache.free_.lower_bound(want);
    if(it != g_hostcache.free_.end() && it->first <= want + want / 4)
    {
      *p = it->second;
      g_hostcache.live[*p] = it->first;
      g_hostcache.cached -= it->first;
      g_hostcache.free_.erase(it);
      return cudaSuccess;
    }
  }
  const cudaError_t e = cudaHostAlloc(p, want, cudaHostAllocDefault);
  if(e == cudaSuccess)
  {
    std::lock_guard<std::mutex> lk(g_hostcache.mu);
    g_hostcache.live[*p] = want;
  }
  return e;
}
cudaError_t host_free(void* p)
{
  if(!p)
    return cudaSuccess;
  std::vector<void*> drop;
  {
    std::lock_guard<std::mutex> lk(g_hostcache.mu);
    auto it = g_hostcache.live.find(p);
    if(it == g_hostcache.live.end())
      drop.push_back(p);
    else
    {
      g_hostcache.free_.insert({it->second, p});
      g_hostcache.cached += it->second;
      g_hostcache.live.erase(it);
      while(g_hostcache.cached > ((size_t)8 << 30) && !g_hostcache.free_.empty())
      {
        auto big = std::prev(g_hostcache.free_.end());
        drop.push_back(big->second);
        g_hostcache.cached -= big->first;
        g_hostcache.free_.erase(big);
      }
    }
  }
  for(void* q : drop)
    cudaFreeHost(q);
  return cudaSuccess;
}
} // namespace
template <class T>
static inline cudaError_t host_alloc_t(T** p, size_t n) { return host_alloc(reinterpret_cast<void**>(p), n); }
/* from here on the engine's device and pinned buffers come from the caches */
#define cudaMalloc(p, n) dev_alloc_t((p), (n))
#define cudaFree(p) dev_free((void*)(p))
#define cudaHostAlloc(p, n, flags) host_alloc_t((p), (n))
#define cudaFreeHost(p) host_free((void*)(p))

/* ---- a plane set: `n` image-shaped 32-bit planes addressed by canvas coordinate ------------- */
struct Planes
{
  int32_t* base = nullptr;
  uint32_t pitch = 0, rows = 0; /* elements, rows */
  uint32_t X0 = 0, Y0 = 0;      /* canvas coordinate stored at column 0 / row 0 */
  int n = 0;
  size_t plane_elems() const { return (size_t)pitch * rows; }
  int32_t* at(int c, uint32_t x, uint32_t y) const { return base + (size_t)c * plane_elems() + (size_t)(y - Y0) * pitch + (x - X0); }
};

static int alloc_planes(Planes& p, int n, uint32_t cx0, uint32_t cy0, uint32_t cx1, uint32_t cy1)
{
  p.n = n;
  p.X0 = cx0 & ~31u; /* 128-byte aligned canvas columns */
  p.Y0 = cy0;
  p.pitch = ((cx1 - p.X0) + 31u + 32u) & ~31u; /* slack: vector loads of halo lanes stay inside */
  p.rows = (cy1 - cy0) + 2;
  CUDA_TRY(cudaMalloc(&p.base, (size_t)n * p.plane_elems() * sizeof(int32_t)));
  CUDA_TRY(cudaMemset(p.base, 0, (size_t)n * p.plane_elems() * sizeof(int32_t)));
  return 0;
}

struct Planes16
{
  uint16_t* base = nullptr;
  uint32_t pitch = 0, rows = 0, X0 = 0, Y0 = 0;
  int n = 0;
  size_t plane_elems() const { return (size_t)pitch * rows; }
  uint16_t* at(int c, uint32_t x, uint32_t y) const { return base + (size_t)c * plane_elems() + (size_t)(y - Y0) * pitch + (x - X0); }
};

struct LevelLaunch
{
  std::vector<DwtLevelDesc> descs;
  DwtLevelDesc* d_descs = nullptr;
  int nc = 1, max_jobs = 0;
  bool point = false;              /* numres = 1: no wavelet level, the launch is the point transform alone */
  uint32_t max_w = 0, max_h = 0;   /* point launches: largest tile component */
  uint64_t alg_bytes = 0; /* one read + one write of every sample of the level */
  std::vector<uint32_t> tile_first; /* descs of selected tile ti are [tile_first[ti], tile_first[ti+1]) */
};

/* Whether the int32 entry points should narrow to 16-bit containers on the host is a property of the
   machine at that moment: it halves the PCIe bytes but triples the host DRAM traffic, so it wins while
   PCIe is the bound (one or two GPUs per socket, or unpinned caller memory) and loses once several
   ranks share a socket's DRAM.  Default policy: time both ways on the first calls, keep the faster,
   look at the other one again every 64 calls. */
/* pipeline chunks per call (tile granular) and their events: purpose 0 chunk done on its stream, 1 chunk's
   upload done, 2 side-stream / scan done, 3 chunk's download done, 4 misc */
#define B2K_MAX_CHUNKS 32
#define CEV(purpose, k) ((purpose) * (B2K_MAX_CHUNKS + 1) + (int)(k))

struct PackTuner
{
  int calls = 0;
  double best[2] = {1e30, 1e30}; /* [0] direct, [1] packed: best wall ms seen while probing */
  double recent = 0;             /* EMA of the chosen mode */
  int choice = -1;
  bool probing_other = false;
  bool next_mode()
  {
    if(choice < 0)
      return (calls & 1) == 0; /* packed, direct, packed, direct, packed, direct */
    probing_other = (calls % 64) == 63;
    return probing_other ? !choice : (choice != 0);
  }
  void record(bool packed, double ms)
  {
    ++calls;
    if(choice < 0)
    {
      if(calls > 1 || !packed) /* the very first packed call allocates the staging buffer */
        best[packed ? 1 : 0] = std::min(best[packed ? 1 : 0], ms);
      if(calls >= 6)
      {
        choice = best[1] < best[0] ? 1 : 0;
        recent = best[choice];
      }
      return;
    }
    if(probing_other)
    {
      if(ms < 0.9 * recent)
      {
        choice = packed ? 1 : 0;
        recent = ms;
      }
      probing_other = false;
      return;
    }
    recent = 0.9 * recent + 0.1 * ms;
  }
};

struct b2k_device_job
{
  b2k_engine* eng = nullptr;
  b2k_coding cp{};
  uint32_t tile_mod = 1, tile_rem = 0;
  TileGrid grid{};
  std::vector<uint32_t> tiles;
  std::vector<Rect> tile_rects;
  std::vector<BandQuant> quant;
  std::vector<b2k_block> blocks;       /* every block, enumeration order */
  std::vector<uint32_t> coded_index;   /* blocks with area, index into `blocks` */
  std::vector<float> dec_quant;        /* per coded block: decoder step / 2^(31-Kmax) */
  std::vector<uint32_t> coded_first;   /* coded blocks of selected tile ti are [coded_first[ti], coded_first[ti+1]) */
  std::vector<uint32_t> chunk_tile;    /* pipeline chunks: selected tiles [chunk_tile[k], chunk_tile[k+1]) */
  cudaEvent_t chunk_ev[5 * (B2K_MAX_CHUNKS + 1)]{}; /* [purpose][chunk], see CEV() */
  uint32_t max_cblk_w = 0;
  HtEncodeLimits enc_limits{0, 0};     /* shared-memory sizing of the HT encoder launches */

  Planes img, coef, ll[2];
  Planes16 img16;                      /* 16-bit sample containers (b2k_encode16 / b2k_decode16), lazily */
  uint16_t* d_ileave = nullptr;        /* pixel-interleaved 16-bit frame (b2k_encode16_interleaved), lazily */
  uint32_t ileave_pitch = 0;           /* in samples: numcomps * width, rounded up to 8 */
  std::vector<LevelLaunch> fwd, inv;   /* launch order */
  std::vector<LevelLaunch> fwd16, inv16; /* finest-level launches re-pointed at img16 (parallel to fwd / inv) */
  HtBlockDesc* d_enc_desc = nullptr;
  HtBlockDesc* d_dec_desc = nullptr;
  std::vector<HtBlockDesc> h_enc_desc;
  HtBlockDesc* h_dec_desc = nullptr;   /* pinned staging */
  HtBlockOut* d_out = nullptr;
  uint64_t* d_offsets = nullptr;
  bool has_crop = false;           /* b2k_decode_window: only this rectangle of the pixels goes back to the host, */
  Rect crop{};                     /* and the host planes are the rectangle's (row 0 / column 0 = crop.y0 / crop.x0) */
  uint32_t* d_recs = nullptr;     /* decode: per-quad records between the two decode phases */
  float* d_dec_quant = nullptr;
  HtBlockOut* d_dec_status = nullptr;
  uint64_t total_quads = 0, group_quads = 0; /* record scratch: closed groups / largest block of the open group */
  uint8_t* d_scratch = nullptr;
  uint64_t scratch_bytes = 0;
  uint8_t* d_bytes = nullptr;
  uint64_t bytes_cap = 0, bytes_used = 0;
  int* d_err = nullptr;
  /* pinned host staging for results */
  HtBlockOut* h_out = nullptr;
  uint64_t* h_offsets = nullptr;
  cudaEvent_t ev[8]{};
  float last_level1_ms = 0.f, last_inv_level1_ms = 0.f;
  uint64_t level1_alg_bytes = 0;
  bool img_is_u16 = false;
  uint16_t* h_stage16 = nullptr;  /* pinned 16-bit staging for the int32 entry points (host_pack.cpp) */
  uint64_t stage16_elems = 0;
  /* ring staging (B2K_STAGE_RING_MB > 0): a few MB-sized pinned slots that are narrowed into / widened out of
     while still cache-resident, instead of one image-sized staging buffer that round-trips through DRAM */
  uint16_t* h_ring = nullptr;
  uint32_t ring_slots = 0;
  uint64_t ring_slot_elems = 0, ring_elems = 0;
  cudaEvent_t ring_ev[16]{};
  PackTuner tune_enc, tune_dec;
  std::vector<cudaEvent_t> q_ev;   /* per-step events of b2k_job_roundtrip_n */
  std::vector<cudaEvent_t> p_ev;   /* per-chunk events of b2k_job_roundtrip_pipelined_n (scan done, chunk done) */
  std::vector<cudaStream_t> p_streams; /* its block-coder streams */
  bool dec_has_refinement = false; /* the block table of the current decode carries SigProp / MagRef passes */
};

/* -------------------------------------------------------------------------------------------- */
extern "C" const char* b2k_last_error(void) { return g_err.c_str(); }
void b2k_set_error(const char* msg) { g_err = msg ? msg : ""; } /* for the other translation units */
extern "C" uint64_t b2k_launch_count(void) { return g_launches.load(); }

extern "C" int32_t b2k_engine_create(int32_t device, b2k_engine** out)
{
  if(!out)
    return -1;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if(e != cudaSuccess || n == 0)
  {
    g_err = std::string("no CUDA device: ") + cudaGetErrorString(e) +
            " (this engine has no CPU path; the host keeps its own)";
    return -1;
  }
  if(device < 0 || device >= n)
  {
    g_err = "device index out of range";
    return -1;
  }
  CUDA_TRY(cudaSetDevice(device));
  b2k_engine* eng = new b2k_engine();
  eng->device = device;
  CUDA_TRY(cudaGetDeviceProperties(&eng->prop, device));
  CUDA_TRY(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&eng->copy_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&eng->h2d_stream, cudaStreamNonBlocking));
  for(cudaStream_t& a : eng->aux)
    CUDA_TRY(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
  *out = eng;
  return 0;
}

extern "C" void b2k_engine_destroy(b2k_engine* e)
{
  if(!e)
    return;
  cudaSetDevice(e->device);
  if(e->cached)
  {
    b2k_job_destroy(e->cached);
    e->cached = nullptr;
  }
  if(e->stream)
    cudaStreamDestroy(e->stream);
  if(e->copy_stream)
    cudaStreamDestroy(e->copy_stream);
  if(e->h2d_stream)
    cudaStreamDestroy(e->h2d_stream);
  for(cudaStream_t a : e->aux)
    if(a)
      cudaStreamDestroy(a);
  delete e;
}

extern "C" void* b2k_host_alloc(size_t bytes)
{
  void* p = nullptr;
  if(cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess)
    return nullptr;
  return p;
}
extern "C" int32_t b2k_set_host_threads(int32_t n)
{
  b2k_host_set_threads(n);
  g_pack_policy.store(n < 0 ? -1 : n == 0 ? 0 : 1);
  return b2k_host_threads();
}

extern "C" int32_t b2k_host_pack_last(int32_t decode) { return g_last_pack[decode ? 1 : 0].load(); }

extern "C" void b2k_host_free(void* p)
{
  if(p)
    cudaFreeHost(p);
}

extern "C" int64_t b2k_enumerate(const b2k_coding* cp, uint32_t tile_mod, uint32_t tile_rem, b2k_block* out, uint64_t cap)
{
  if(!cp || tile_mod == 0)
    return -1;
  if(const char* why = unsupported_reason(*cp))
  {
    g_err = why;
    return -1;
  }
  const TileGrid g = tile_grid(*cp);
  const std::vector<BandQuant> q = band_quant(*cp);
  std::vector<b2k_block> v;
  for(uint32_t t = 0; t < g.nx * g.ny; ++t)
    if(t % tile_mod == tile_rem)
      enumerate_tile_blocks(*cp, t, tile_rect(*cp, g, t), q, v);
  for(uint64_t i = 0; i < v.size() && i < cap; ++i)
    out[i] = v[i];
  return (int64_t)v.size();
}

/* -------------------------------------------------------------------------------------------- */
static void fill_strips(DwtLevelDesc& d, int pairs_per_seg)
{
  const int span = d.u1 - (d.u0 & ~7);
  const int nstrips = std::max(1, (span + 239) / 240);
  int sw = (span + nstrips - 1) / nstrips;
  sw = (sw + 7) & ~7;
  d.nstrips = (uint16_t)nstrips;
  d.strip_w = (uint16_t)sw;
  const int npairs = ((d.v1 - 1) >> 1) - (d.v0 >> 1) + 1;
  d.pairs_per_seg = (uint16_t)pairs_per_seg;
  d.nsegs = (uint16_t)((npairs + pairs_per_seg - 1) / pairs_per_seg);
}

static int upload_descs(LevelLaunch& L)
{
  L.max_jobs = 0;
  for(const DwtLevelDesc& d : L.descs)
    L.max_jobs = std::max(L.max_jobs, (int)d.nstrips * (int)d.nsegs);
  if(L.descs.empty())
    return 0;
  CUDA_TRY(cudaMalloc(&L.d_descs, L.descs.size() * sizeof(DwtLevelDesc)));
  CUDA_TRY(cudaMemcpy(L.d_descs, L.descs.data(), L.descs.size() * sizeof(DwtLevelDesc), cudaMemcpyHostToDevice));
  return 0;
}

static int build_dwt_plan(b2k_device_job* J)
{
  const b2k_coding& cp = J->cp;
  const int L = cp.numres - 1;
  const int ncomp = cp.numcomps;
  const int pairs53 = 32, pairs97 = 32;
  const int P = cp.irreversible ? pairs97 : pairs53;
  const int32_t dc = cp.sgnd ? 0 : -(1 << (cp.prec - 1));
  const int32_t lo = cp.sgnd ? -(1 << (cp.prec - 1)) : 0, hi = cp.sgnd ? (1 << (cp.prec - 1)) - 1 : (1 << cp.prec) - 1;

  if(L == 0)
  { /* one resolution: the tile is its own LL band; what is left is DC shift + colour transform (dwt.cu k_point_transform) */
    for(int dir = 0; dir < 2; ++dir)
    {
      std::vector<LevelLaunch>& out = dir == 0 ? J->fwd : J->inv;
      LevelLaunch mctL, sglL;
      mctL.nc = 3;
      sglL.nc = 1;
      mctL.point = sglL.point = true;
      for(size_t ti = 0; ti < J->tiles.size(); ++ti)
      {
        mctL.tile_first.push_back((uint32_t)mctL.descs.size());
        sglL.tile_first.push_back((uint32_t)sglL.descs.size());
        const Rect tc = J->tile_rects[ti];
        if(tc.empty())
          continue;
        for(int c = 0; c < ncomp;)
        {
          const bool group = cp.mct && c == 0;
          const int nc = group ? 3 : 1;
          LevelLaunch& LL = group ? mctL : sglL;
          DwtLevelDesc d{};
          d.u0 = (int32_t)tc.x0; d.v0 = (int32_t)tc.y0; d.u1 = (int32_t)tc.x1; d.v1 = (int32_t)tc.y1;
          d.first_level = 1;
          d.comp0 = (uint8_t)c;
          for(int k = 0; k < nc; ++k)
          {
            d.in[k] = J->img.at(c + k, tc.x0, tc.y0);
            d.in_pitch = J->img.pitch;
            d.out_c[k] = J->coef.at(c + k, tc.x0, tc.y0);
            d.out_ll[k] = d.out_c[k];
            d.c_pitch = d.ll_pitch = J->coef.pitch;
            d.shift[k] = dc;
            d.lo[k] = lo;
            d.hi[k] = hi;
          }
          LL.descs.push_back(d);
          LL.max_w = std::max(LL.max_w, tc.w());
          LL.max_h = std::max(LL.max_h, tc.h());
          LL.alg_bytes += (uint64_t)tc.w() * tc.h() * nc * 8;
          c += nc;
        }
      }
      mctL.tile_first.push_back((uint32_t)mctL.descs.size());
      sglL.tile_first.push_back((uint32_t)sglL.descs.size());
      if(!mctL.descs.empty()) out.push_back(std::move(mctL));
      if(!sglL.descs.empty()) out.push_back(std::move(sglL));
      for(LevelLaunch& Lh : out)
      {
        Lh.max_jobs = 1;
        if(upload_descs(Lh))
          return -1;
      }
    }
    return 0;
  }
  /* forward: level 1 (MCT group, then the rest), then levels 2..L component-wise */
  for(int dir = 0; dir < 2; ++dir)
  {
    std::vector<LevelLaunch>& out = dir == 0 ? J->fwd : J->inv;
    for(int lvl = 1; lvl <= L; ++lvl)
    {
      const int resno = cp.numres - lvl; /* resolution being split / rebuilt */
      LevelLaunch mctL, sglL;
      mctL.nc = 3;
      sglL.nc = 1;
      for(size_t ti = 0; ti < J->tiles.size(); ++ti)
      {
        mctL.tile_first.push_back((uint32_t)mctL.descs.size());
        sglL.tile_first.push_back((uint32_t)sglL.descs.size());
        const Rect tc = J->tile_rects[ti];
        const Rect r = resolution_rect(tc, cp.numres, resno);
        if(r.empty())
          continue;
        for(int c = 0; c < ncomp;)
        {
          const bool group = (lvl == 1 && cp.mct && c == 0);
          const int nc = group ? 3 : 1;
          DwtLevelDesc d{};
          d.u0 = (int32_t)r.x0; d.v0 = (int32_t)r.y0; d.u1 = (int32_t)r.x1; d.v1 = (int32_t)r.y1;
          d.first_level = lvl == 1;
          d.in_is_u16 = 0;
          d.comp0 = (uint8_t)c;
          const uint32_t llx = (r.x0 + 1) >> 1, lly = (r.y0 + 1) >> 1;
          for(int k = 0; k < nc; ++k)
          {
            const int cc = c + k;
            /* finer side: image at level 1, else LL scratch written by level lvl-1 */
            const Planes& fine = (lvl == 1) ? J->img : J->ll[(lvl - 1) & 1];
            d.in[k] = fine.at(cc, r.x0, r.y0);
            d.in_pitch = fine.pitch;
            d.out_c[k] = J->coef.at(cc, tc.x0, tc.y0);
            d.c_pitch = J->coef.pitch;
            if(lvl == L)
            {
              d.out_ll[k] = d.out_c[k];
              d.ll_pitch = J->coef.pitch;
            }
            else
            {
              const Planes& coarse = J->ll[lvl & 1];
              d.out_ll[k] = coarse.at(cc, llx, lly);
              d.ll_pitch = coarse.pitch;
            }
            d.shift[k] = dc;
            d.lo[k] = lo;
            d.hi[k] = hi;
          }
          fill_strips(d, P);
          (group ? mctL : sglL).descs.push_back(d);
          const uint64_t samples = (uint64_t)r.w() * r.h() * nc;
          (group ? mctL : sglL).alg_bytes += samples * 8;
          c += nc;
        }
      }
      mctL.tile_first.push_back((uint32_t)mctL.descs.size());
      sglL.tile_first.push_back((uint32_t)sglL.descs.size());
      /* Coarse levels have few rows: with 32 row pairs per warp-job the whole level is a handful of long serial
         chains on an almost empty GPU (level 3 of config 2: 34.7 us for 64 MB).  Cut the segments until the level
         offers about two warp-jobs per resident warp (or segments of 8 pairs, where the recomputed halo rows start
         to dominate): the work is L2-resident there, parallelism is what it lacks. */
      for(LevelLaunch* LL : {&mctL, &sglL})
      {
        int Pl = P;
        auto jobs = [&] {
          uint64_t n = 0;
          for(const DwtLevelDesc& d : LL->descs)
            n += (uint64_t)d.nstrips * d.nsegs;
          return n;
        };
        while(!LL->descs.empty() && Pl > 8 && jobs() < 3552)
        {
          Pl >>= 1;
          for(DwtLevelDesc& d : LL->descs)
            fill_strips(d, Pl);
        }
      }
      if(dir == 0)
      {
        if(!mctL.descs.empty()) out.push_back(std::move(mctL));
        if(!sglL.descs.empty()) out.push_back(std::move(sglL));
      }
      else
      {
        if(!sglL.descs.empty()) out.insert(out.begin(), std::move(sglL));
        if(!mctL.descs.empty()) out.insert(out.begin(), std::move(mctL));
      }
    }
    for(LevelLaunch& Lh : out)
      if(upload_descs(Lh))
        return -1;
  }
  return 0;
}

static uint32_t slot_capacity(uint32_t w, uint32_t h, uint32_t kmax)
{
  /* MagSgn: <= (kmax+2) bits per sample, 8/7 stuffing; VLC: <= 15 bits per quad, 8/7; MEL 256 */
  const uint64_t samples = (uint64_t)w * h, quads = (uint64_t)((w + 1) / 2) * ((h + 1) / 2);
  const uint64_t ms = (samples * (kmax + 2) + 6) / 7 + 16;
  const uint64_t vlc = (quads * 15 + 6) / 7 + 16;
  return (uint32_t)((ms + vlc + 256 + 15) & ~15ull);
}

static int build_block_plan(b2k_device_job* J)
{
  const b2k_coding& cp = J->cp;
  J->blocks.clear();
  for(size_t ti = 0; ti < J->tiles.size(); ++ti)
    enumerate_tile_blocks(cp, J->tiles[ti], J->tile_rects[ti], J->quant, J->blocks);
  /* map tile index -> rect */
  std::vector<Rect> rect_of(J->grid.nx * J->grid.ny);
  for(size_t ti = 0; ti < J->tiles.size(); ++ti)
    rect_of[J->tiles[ti]] = J->tile_rects[ti];
  uint64_t off = 0;
  std::vector<uint32_t> sel_of(J->grid.nx * J->grid.ny, 0);
  for(size_t ti = 0; ti < J->tiles.size(); ++ti)
    sel_of[J->tiles[ti]] = (uint32_t)ti;
  J->coded_first.assign(J->tiles.size() + 1, 0);
  for(uint32_t i = 0; i < J->blocks.size(); ++i)
  {
    const b2k_block& b = J->blocks[i];
    if(b.x1 <= b.x0 || b.y1 <= b.y0)
      continue;
    J->coded_first[sel_of[b.tile] + 1] = (uint32_t)J->h_enc_desc.size() + 1;
    const Rect& tr = rect_of[b.tile];
    const BandQuant& bq = J->quant[band_quant_index(b.resno, b.orient)];
    HtBlockDesc d{};
    d.coef = J->coef.at(b.comp, tr.x0 + b.buf_x, tr.y0 + b.buf_y);
    d.pitch = J->coef.pitch;
    d.w = (uint16_t)(b.x1 - b.x0);
    d.h = (uint16_t)(b.y1 - b.y0);
    d.kmax = bq.kmax;
    d.irreversible = cp.irreversible;
    d.quant = 1.0f / bq.step_enc; /* CompressScheduler.cpp L131: inv_step_ht */
    d.slot_cap = slot_capacity(d.w, d.h, bq.kmax);
    d.slot_off = off;
    off += d.slot_cap;
    { /* per-quad records of 32 consecutive coded blocks are interleaved word by word (ht_dec.cu phase A): a group of 32
         takes 32 x its largest block's quads; entry k of block j of the group sits at group_base + 32 k + j */
      const size_t j = J->h_enc_desc.size() & 31u;
      const uint64_t quads = (uint64_t)((d.w + 1) / 2) * ((d.h + 1) / 2);
      if(j == 0)
      {
        J->total_quads += 32 * J->group_quads; /* close the previous group */
        J->group_quads = 0;
      }
      J->group_quads = std::max(J->group_quads, quads);
      d.rec_off = (uint32_t)(J->total_quads + j);
    }
    J->max_cblk_w = std::max<uint32_t>(J->max_cblk_w, d.w);
    J->enc_limits.stage_words = std::max(J->enc_limits.stage_words, b2k_ht_encode_stage_words(d.w));
    J->enc_limits.max_kmax = std::max<uint32_t>(J->enc_limits.max_kmax, d.kmax);
    J->h_enc_desc.push_back(d);
    J->dec_quant.push_back(bq.step_dec / (float)(1u << (31 - bq.kmax)));
    J->coded_index.push_back(i);
  }
  J->scratch_bytes = off;
  for(size_t ti = 1; ti <= J->tiles.size(); ++ti) /* tiles without coded blocks inherit the running count */
    J->coded_first[ti] = std::max(J->coded_first[ti], J->coded_first[ti - 1]);
  {
    const uint32_t nt = (uint32_t)J->tiles.size();
    uint32_t want = 8;
    if(const char* ev = getenv("B2K_CHUNKS"))
      want = (uint32_t)std::max(1, std::min(B2K_MAX_CHUNKS, atoi(ev)));
    const uint32_t nchunks = std::max(1u, std::min(want, nt));
    J->chunk_tile.clear();
    for(uint32_t k = 0; k <= nchunks; ++k)
      J->chunk_tile.push_back((uint32_t)((uint64_t)k * nt / nchunks));
  }
  const size_t n = J->h_enc_desc.size();
  if(n)
  {
    CUDA_TRY(cudaMalloc(&J->d_enc_desc, n * sizeof(HtBlockDesc)));
    CUDA_TRY(cudaMemcpy(J->d_enc_desc, J->h_enc_desc.data(), n * sizeof(HtBlockDesc), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&J->d_dec_desc, n * sizeof(HtBlockDesc)));
    CUDA_TRY(cudaMalloc(&J->d_out, n * sizeof(HtBlockOut)));
    CUDA_TRY(cudaMalloc(&J->d_offsets, (n + 1) * sizeof(uint64_t)));
    CUDA_TRY(cudaMemset(J->d_offsets, 0, (n + 1) * sizeof(uint64_t))); /* offsets[0] stays 0: the scan's base */
    CUDA_TRY(cudaMalloc(&J->d_scratch, J->scratch_bytes + 64));
    CUDA_TRY(cudaMalloc(&J->d_recs, (J->total_quads + 32 * J->group_quads + 64) * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&J->d_dec_status, n * sizeof(HtBlockOut)));
    CUDA_TRY(cudaMalloc(&J->d_dec_quant, n * sizeof(float)));
    CUDA_TRY(cudaMemcpy(J->d_dec_quant, J->dec_quant.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaHostAlloc(&J->h_out, n * sizeof(HtBlockOut), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&J->h_dec_desc, n * sizeof(HtBlockDesc), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&J->h_offsets, (n + 1) * sizeof(uint64_t), cudaHostAllocDefault));
  }
  CUDA_TRY(cudaMalloc(&J->d_err, sizeof(int)));
  CUDA_TRY(cudaMemset(J->d_err, 0, sizeof(int)));
  return 0;
}

extern "C" int32_t b2k_job_create(b2k_engine* e, const b2k_coding* cp, uint32_t tile_mod, uint32_t tile_rem,
                                  b2k_device_job** out)
{
  if(!e || !cp || !out || tile_mod == 0)
    return -1;
  *out = nullptr;
  if(const char* why = unsupported_reason(*cp))
  {
    g_err = why;
    return 1; /* not handled: the host keeps its CPU path (plugin_accelerate.h L32-36) */
  }
  CUDA_TRY(cudaSetDevice(e->device));
  b2k_device_job* J = new b2k_device_job();
  J->eng = e;
  J->cp = *cp;
  J->tile_mod = tile_mod;
  J->tile_rem = tile_rem;
  J->grid = tile_grid(*cp);
  J->quant = band_quant(*cp);
  for(uint32_t t = 0; t < J->grid.nx * J->grid.ny; ++t)
    if(t % tile_mod == tile_rem)
    {
      const Rect r = tile_rect(*cp, J->grid, t);
      if(r.empty())
        continue;
      J->tiles.push_back(t);
      J->tile_rects.push_back(r);
    }
  const int nc = cp->numcomps;
  if(alloc_planes(J->img, nc, cp->x0, cp->y0, cp->x1, cp->y1) || alloc_planes(J->coef, nc, cp->x0, cp->y0, cp->x1, cp->y1))
  {
    b2k_job_destroy(J);
    return -1;
  }
  /* LL scratch, canvas origin (0,0): ll[1] receives the LL of odd levels (1/2, 1/8, ... resolution),
     ll[0] of even levels (1/4, 1/16, ...); a level reads one and writes the other */
  if(alloc_planes(J->ll[1], nc, 0, 0, ((cp->x1 + 1) >> 1) + 1, ((cp->y1 + 1) >> 1) + 1) ||
     alloc_planes(J->ll[0], nc, 0, 0, ((cp->x1 + 3) >> 2) + 1, ((cp->y1 + 3) >> 2) + 1))
  {
    b2k_job_destroy(J);
    return -1;
  }
  if(build_dwt_plan(J) || build_block_plan(J))
  {
    b2k_job_destroy(J);
    return -1;
  }
  for(cudaEvent_t& ev : J->ev)
    CUDA_TRY(cudaEventCreate(&ev));
  for(cudaEvent_t& ev : J->chunk_ev)
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  *out = J;
  return 0;
}

extern "C" void b2k_job_destroy(b2k_device_job* J)
{
  if(!J)
    return;
  cudaSetDevice(J->eng->device);
  cudaStreamSynchronize(J->eng->stream);
  cudaFree(J->img.base);
  cudaFree(J->img16.base);
  cudaFree(J->d_ileave);
  for(LevelLaunch& L : J->fwd16) cudaFree(L.d_descs);
  for(LevelLaunch& L : J->inv16) cudaFree(L.d_descs);
  cudaFree(J->coef.base);
  cudaFree(J->ll[0].base);
  cudaFree(J->ll[1].base);
  for(LevelLaunch& L : J->fwd) cudaFree(L.d_descs);
  for(LevelLaunch& L : J->inv) cudaFree(L.d_descs);
  cudaFree(J->d_enc_desc);
  cudaFree(J->d_dec_desc);
  cudaFree(J->d_out);
  cudaFree(J->d_offsets);
  cudaFree(J->d_scratch);
  cudaFree(J->d_recs);
  cudaFree(J->d_dec_quant);
  cudaFree(J->d_dec_status);
  cudaFree(J->d_bytes);
  cudaFree(J->d_err);
  cudaFreeHost(J->h_out);
  cudaFreeHost(J->h_dec_desc);
  cudaFreeHost(J->h_offsets);
  cudaFreeHost(J->h_stage16);
  cudaFreeHost(J->h_ring);
  for(cudaEvent_t ev : J->q_ev)
    cudaEventDestroy(ev);
  for(cudaEvent_t ev : J->p_ev)
    cudaEventDestroy(ev);
  for(cudaStream_t ps : J->p_streams)
    cudaStreamDestroy(ps);
  for(cudaEvent_t& ev : J->ring_ev)
    if(ev)
      cudaEventDestroy(ev);
  for(cudaEvent_t& ev : J->ev)
    if(ev)
      cudaEventDestroy(ev);
  for(cudaEvent_t& ev : J->chunk_ev)
    if(ev)
      cudaEventDestroy(ev);
  delete J;
}

extern "C" uint64_t b2k_job_num_blocks(const b2k_device_job* J) { return J ? J->blocks.size() : 0; }

/* ---- host <-> device plane copies, per selected tile ---------------------------------------- */
static int copy_planes(b2k_device_job* J, const Planes& P, void* const* host, const uint32_t* strides, bool to_device,
                       cudaStream_t st, size_t t0 = 0, size_t t1 = (size_t)-1)
{
  const b2k_coding& cp = J->cp;
  t1 = std::min(t1, J->tiles.size());
  for(size_t ti = t0; ti < t1;)
  {
    /* merge horizontally adjacent selected tiles of one tile row into a single rectangle */
    Rect r = J->tile_rects[ti];
    size_t tj = ti + 1;
    while(tj < t1 && J->tile_rects[tj].y0 == r.y0 && J->tile_rects[tj].y1 == r.y1 && J->tile_rects[tj].x0 == r.x1)
    {
      r.x1 = J->tile_rects[tj].x1;
      ++tj;
    }
    uint32_t ox = cp.x0, oy = cp.y0; /* canvas position of the host planes' first sample */
    if(!to_device && J->has_crop)
    {
      r.x0 = std::max(r.x0, J->crop.x0); r.y0 = std::max(r.y0, J->crop.y0);
      r.x1 = std::min(r.x1, J->crop.x1); r.y1 = std::min(r.y1, J->crop.y1);
      ox = J->crop.x0; oy = J->crop.y0;
      if(r.x1 <= r.x0 || r.y1 <= r.y0)
      {
        ti = tj;
        continue;
      }
    }
    for(int c = 0; c < cp.numcomps; ++c)
    {
      int32_t* dev = P.at(c, r.x0, r.y0);
      int32_t* hst = reinterpret_cast<int32_t*>(host[c]) + (size_t)(r.y0 - oy) * strides[c] + (r.x0 - ox);
      if(to_device)
        CUDA_TRY(cudaMemcpy2DAsync(dev, (size_t)P.pitch * 4, hst, (size_t)strides[c] * 4, (size_t)r.w() * 4, r.h(),
                                   cudaMemcpyHostToDevice, st));
      else
        CUDA_TRY(cudaMemcpy2DAsync(hst, (size_t)strides[c] * 4, dev, (size_t)P.pitch * 4, (size_t)r.w() * 4, r.h(),
                                   cudaMemcpyDeviceToHost, st));
    }
    ti = tj;
  }
  return 0;
}

extern "C" int32_t b2k_job_upload(b2k_device_job* J, const int32_t* const* planes, const uint32_t* strides)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  if(copy_planes(J, J->img, (void* const*)planes, strides, true, J->eng->stream)) return -1;
  CUDA_TRY(cudaStreamSynchronize(J->eng->stream));
  return 0;
}
extern "C" int32_t b2k_job_download(b2k_device_job* J, int32_t* const* planes, const uint32_t* strides)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  if(copy_planes(J, J->img, (void* const*)planes, strides, false, J->eng->stream)) return -1;
  CUDA_TRY(cudaStreamSynchronize(J->eng->stream));
  return 0;
}
extern "C" int32_t b2k_job_download_coeffs(b2k_device_job* J, int32_t* const* planes, const uint32_t* strides)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  if(copy_planes(J, J->coef, (void* const*)planes, strides, false, J->eng->stream)) return -1;
  CUDA_TRY(cudaStreamSynchronize(J->eng->stream));
  return 0;
}
extern "C" int32_t b2k_job_upload_coeffs(b2k_device_job* J, const int32_t* const* planes, const uint32_t* strides)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  if(copy_planes(J, J->coef, (void* const*)planes, strides, true, J->eng->stream)) return -1;
  CUDA_TRY(cudaStreamSynchronize(J->eng->stream));
  return 0;
}


/* ---- 16-bit sample containers ---------------------------------------------------------------- */
static int ensure_u16(b2k_device_job* J)
{
  if(J->img16.base)
    return 0;
  const b2k_coding& cp = J->cp;
  Planes16& P = J->img16;
  P.n = cp.numcomps;
  P.X0 = cp.x0 & ~63u;
  P.Y0 = cp.y0;
  P.pitch = ((cp.x1 - P.X0) + 7u) & ~7u; /* no slack: a full-width tile row is one contiguous block */
  P.rows = (cp.y1 - cp.y0) + 2;
  CUDA_TRY(cudaMalloc(&P.base, (size_t)P.n * P.plane_elems() * sizeof(uint16_t)));
  CUDA_TRY(cudaMemset(P.base, 0, (size_t)P.n * P.plane_elems() * sizeof(uint16_t)));
  return 0;
}

/* widen (after H2D) or narrow (before D2H) the rectangles of selected tiles [t0, t1) */
static int convert_planes16(b2k_device_job* J, bool widen, cudaStream_t st, size_t t0, size_t t1)
{
  const b2k_coding& cp = J->cp;
  t1 = std::min(t1, J->tiles.size());
  for(size_t ti = t0; ti < t1;)
  {
    Rect r = J->tile_rects[ti];
    size_t tj = ti + 1;
    while(tj < t1 && J->tile_rects[tj].y0 == r.y0 && J->tile_rects[tj].y1 == r.y1 && J->tile_rects[tj].x0 == r.x1)
    {
      r.x1 = J->tile_rects[tj].x1;
      ++tj;
    }
    for(int c = 0; c < cp.numcomps; ++c)
    {
      if(widen)
        b2k_launch_widen16(J->img16.at(c, r.x0, r.y0), J->img16.pitch, J->img.at(c, r.x0, r.y0), J->img.pitch, r.w(), r.h(),
                           cp.sgnd, st);
      else
        b2k_launch_narrow16(J->img.at(c, r.x0, r.y0), J->img.pitch, J->img16.at(c, r.x0, r.y0), J->img16.pitch, r.w(), r.h(), st);
    }
    ti = tj;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int copy_planes16(b2k_device_job* J, void* const* host, const uint32_t* strides, bool to_device, cudaStream_t st,
                         size_t t0, size_t t1)
{
  const b2k_coding& cp = J->cp;
  const Planes16& P = J->img16;
  t1 = std::min(t1, J->tiles.size());
  for(size_t ti = t0; ti < t1;)
  {
    Rect r = J->tile_rects[ti];
    size_t tj = ti + 1;
    while(tj < t1 && J->tile_rects[tj].y0 == r.y0 && J->tile_rects[tj].y1 == r.y1 && J->tile_rects[tj].x0 == r.x1)
    {
      r.x1 = J->tile_rects[tj].x1;
      ++tj;
    }
    uint32_t ox = cp.x0, oy = cp.y0;
    if(!to_device && J->has_crop)
    {
      r.x0 = std::max(r.x0, J->crop.x0); r.y0 = std::max(r.y0, J->crop.y0);
      r.x1 = std::min(r.x1, J->crop.x1); r.y1 = std::min(r.y1, J->crop.y1);
      ox = J->crop.x0; oy = J->crop.y0;
      if(r.x1 <= r.x0 || r.y1 <= r.y0)
      {
        ti = tj;
        continue;
      }
    }
    for(int c = 0; c < cp.numcomps; ++c)
    {
      uint16_t* dev = P.at(c, r.x0, r.y0);
      uint16_t* hst = reinterpret_cast<uint16_t*>(host[c]) + (size_t)(r.y0 - oy) * strides[c] + (r.x0 - ox);
      if(strides[c] == P.pitch && r.w() == P.pitch)
      { /* contiguous on both sides: one linear copy */
        if(to_device)
          CUDA_TRY(cudaMemcpyAsync(dev, hst, (size_t)r.w() * r.h() * 2, cudaMemcpyHostToDevice, st));
        else
          CUDA_TRY(cudaMemcpyAsync(hst, dev, (size_t)r.w() * r.h() * 2, cudaMemcpyDeviceToHost, st));
        continue;
      }
      if(to_device)
        CUDA_TRY(cudaMemcpy2DAsync(dev, (size_t)P.pitch * 2, hst, (size_t)strides[c] * 2, (size_t)r.w() * 2, r.h(),
                                   cudaMemcpyHostToDevice, st));
      else
        CUDA_TRY(cudaMemcpy2DAsync(hst, (size_t)strides[c] * 2, dev, (size_t)P.pitch * 2, (size_t)r.w() * 2, r.h(),
                                   cudaMemcpyDeviceToHost, st));
    }
    ti = tj;
  }
  return 0;
}

/* ---- pixel-interleaved 16-bit frames: the rows cross PCIe as they are, the planes are made on the device ---- */
static int ensure_ileave(b2k_device_job* J)
{
  if(J->d_ileave)
    return 0;
  const b2k_coding& cp = J->cp;
  J->ileave_pitch = ((cp.x1 - cp.x0) * cp.numcomps + 7u) & ~7u;
  CUDA_TRY(cudaMalloc(&J->d_ileave, (size_t)J->ileave_pitch * (cp.y1 - cp.y0) * sizeof(uint16_t)));
  return 0;
}

template <typename F>
static void for_tile_row_runs(const b2k_device_job* J, size_t t0, size_t t1, F&& fn)
{
  t1 = std::min(t1, J->tiles.size());
  for(size_t ti = t0; ti < t1;)
  {
    Rect r = J->tile_rects[ti];
    size_t tj = ti + 1;
    while(tj < t1 && J->tile_rects[tj].y0 == r.y0 && J->tile_rects[tj].y1 == r.y1 && J->tile_rects[tj].x0 == r.x1)
    {
      r.x1 = J->tile_rects[tj].x1;
      ++tj;
    }
    fn(r);
    ti = tj;
  }
}

static int upload_interleaved(b2k_device_job* J, const uint16_t* host, uint32_t stride, cudaStream_t st, size_t t0, size_t t1)
{
  const b2k_coding& cp = J->cp;
  const uint32_t nc = cp.numcomps;
  cudaError_t err = cudaSuccess;
  for_tile_row_runs(J, t0, t1, [&](const Rect& r) {
    uint16_t* dev = J->d_ileave + (size_t)(r.y0 - cp.y0) * J->ileave_pitch + (size_t)(r.x0 - cp.x0) * nc;
    const uint16_t* hst = host + (size_t)(r.y0 - cp.y0) * stride + (size_t)(r.x0 - cp.x0) * nc;
    const size_t row_bytes = (size_t)r.w() * nc * 2;
    cudaError_t e = (stride == J->ileave_pitch && r.w() == cp.x1 - cp.x0)
                        ? cudaMemcpyAsync(dev, hst, (size_t)stride * 2 * (r.h() - 1) + row_bytes, cudaMemcpyHostToDevice, st)
                        : cudaMemcpy2DAsync(dev, (size_t)J->ileave_pitch * 2, hst, (size_t)stride * 2, row_bytes, r.h(),
                                            cudaMemcpyHostToDevice, st);
    if(e != cudaSuccess)
      err = e;
  });
  CUDA_TRY(err);
  return 0;
}

static int split_interleaved(b2k_device_job* J, cudaStream_t st, size_t t0, size_t t1)
{
  const b2k_coding& cp = J->cp;
  for_tile_row_runs(J, t0, t1, [&](const Rect& r) {
    int32_t* dst[4] = {nullptr, nullptr, nullptr, nullptr};
    for(int c = 0; c < cp.numcomps; ++c)
      dst[c] = J->img.at(c, r.x0, r.y0);
    b2k_launch_widen16_interleaved(J->d_ileave + (size_t)(r.y0 - cp.y0) * J->ileave_pitch + (size_t)(r.x0 - cp.x0) * cp.numcomps,
                                   J->ileave_pitch, dst, cp.numcomps, J->img.pitch, r.w(), r.h(), cp.sgnd, st);
  });
  CUDA_TRY(cudaGetLastError());
  return 0;
}

/* int32 entry points with samples of <= 16 bits: the PCIe legs carry 16-bit containers, converted
   chunk by chunk on host threads (host_pack.cpp) while the neighbouring chunk is on the bus */
static bool host_pack_eligible(const b2k_device_job* J)
{
  const b2k_coding& cp = J->cp;
  const uint64_t samples = (uint64_t)(cp.x1 - cp.x0) * (cp.y1 - cp.y0) * cp.numcomps;
  return cp.prec <= 16 && samples >= (1u << 22) && J->chunk_tile.size() > 2 && b2k_host_threads() > 0;
}

static int ensure_stage16(b2k_device_job* J)
{
  if(J->h_stage16)
    return 0;
  const b2k_coding& cp = J->cp;
  J->stage16_elems = (uint64_t)(cp.x1 - cp.x0) * (cp.y1 - cp.y0);
  CUDA_TRY(cudaHostAlloc(&J->h_stage16, J->stage16_elems * cp.numcomps * sizeof(uint16_t), cudaHostAllocDefault));
  return 0;
}
/* staging planes are image-shaped, stride = image width */
static void stage16_views(const b2k_device_job* J, void** planes, uint32_t* strides)
{
  for(int c = 0; c < J->cp.numcomps; ++c)
  {
    planes[c] = J->h_stage16 + (uint64_t)c * J->stage16_elems;
    strides[c] = J->cp.x1 - J->cp.x0;
  }
}
static void host_convert_chunk(const b2k_device_job* J, void* const* user, const uint32_t* strides, bool widen, size_t t0,
                               size_t t1)
{
  const b2k_coding& cp = J->cp;
  const uint32_t W = cp.x1 - cp.x0;
  std::vector<b2k_host_rect> rects;
  t1 = std::min(t1, J->tiles.size());
  for(size_t ti = t0; ti < t1;)
  {
    Rect r = J->tile_rects[ti];
    size_t tj = ti + 1;
    while(tj < t1 && J->tile_rects[tj].y0 == r.y0 && J->tile_rects[tj].y1 == r.y1 && J->tile_rects[tj].x0 == r.x1)
    {
      r.x1 = J->tile_rects[tj].x1;
      ++tj;
    }
    for(int c = 0; c < cp.numcomps; ++c)
    {
      int32_t* u = reinterpret_cast<int32_t*>(user[c]) + (size_t)(r.y0 - cp.y0) * strides[c] + (r.x0 - cp.x0);
      uint16_t* s = J->h_stage16 + (uint64_t)c * J->stage16_elems + (size_t)(r.y0 - cp.y0) * W + (r.x0 - cp.x0);
      if(widen)
        rects.push_back({s, u, W, strides[c], r.w(), r.h()});
      else
        rects.push_back({u, s, strides[c], W, r.w(), r.h()});
    }
    ti = tj;
  }
  static const bool dbg = getenv("B2K_DEBUG_TIMING") != nullptr;
  const auto t_a = std::chrono::steady_clock::now();
  b2k_host_convert(rects.data(), rects.size(), widen, cp.sgnd != 0);
  if(dbg)
    fprintf(stderr, "[b2k] host %s tiles [%zu,%zu): %.3f ms\n", widen ? "widen" : "narrow", t0, t1,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a).count());
}

/* ---- ring staging ---------------------------------------------------------------------------- */
struct StagePiece
{
  uint32_t chunk, comp, x0, y0, w, rows; /* canvas coordinates */
};
struct RingGeom
{
  uint32_t slot_mb, slots;
};
/* B2K_RING_ENC / B2K_RING_DEC = "<slot MB>,<slots>" ("0" = image-sized staging instead of a ring) */
static RingGeom ring_geom(bool decode)
{
  static const RingGeom g[2] = {[] {
                                  RingGeom r{8, 4};
                                  if(const char* e = getenv("B2K_RING_ENC"))
                                  {
                                    r.slot_mb = (uint32_t)std::max(0, atoi(e));
                                    if(const char* c = strchr(e, ','))
                                      r.slots = (uint32_t)std::max(2, std::min(16, atoi(c + 1)));
                                  }
                                  return r;
                                }(),
                                [] {
                                  RingGeom r{16, 4};
                                  if(const char* e = getenv("B2K_RING_DEC"))
                                  {
                                    r.slot_mb = (uint32_t)std::max(0, atoi(e));
                                    if(const char* c = strchr(e, ','))
                                      r.slots = (uint32_t)std::max(2, std::min(16, atoi(c + 1)));
                                  }
                                  return r;
                                }()};
  return g[decode ? 1 : 0];
}
static int ensure_ring(b2k_device_job* J, bool decode)
{
  const RingGeom g = ring_geom(decode);
  const uint64_t min_elems = (uint64_t)(J->cp.x1 - J->cp.x0) * 4;
  const uint64_t slot_elems = std::max<uint64_t>((uint64_t)g.slot_mb * (1u << 20) / sizeof(uint16_t), min_elems);
  const uint64_t need = slot_elems * g.slots;
  if(need > J->ring_elems)
  {
    if(J->h_ring)
      cudaFreeHost(J->h_ring);
    J->h_ring = nullptr;
    CUDA_TRY(cudaHostAlloc(&J->h_ring, need * sizeof(uint16_t), cudaHostAllocDefault));
    J->ring_elems = need;
  }
  for(uint32_t i = 0; i < g.slots; ++i)
    if(!J->ring_ev[i])
      CUDA_TRY(cudaEventCreateWithFlags(&J->ring_ev[i], cudaEventDisableTiming));
  J->ring_slots = g.slots;
  J->ring_slot_elems = slot_elems;
  return 0;
}
static void chunk_pieces(const b2k_device_job* J, uint32_t chunk, std::vector<StagePiece>& out)
{
  const b2k_coding& cp = J->cp;
  const size_t t0 = J->chunk_tile[chunk], t1 = std::min<size_t>(J->chunk_tile[chunk + 1], J->tiles.size());
  for(size_t ti = t0; ti < t1;)
  {
    Rect r = J->tile_rects[ti];
    size_t tj = ti + 1;
    while(tj < t1 && J->tile_rects[tj].y0 == r.y0 && J->tile_rects[tj].y1 == r.y1 && J->tile_rects[tj].x0 == r.x1)
    {
      r.x1 = J->tile_rects[tj].x1;
      ++tj;
    }
    const uint32_t rows_per = (uint32_t)std::max<uint64_t>(1, J->ring_slot_elems / r.w());
    for(uint32_t y = r.y0; y < r.y1; y += rows_per)
      for(uint32_t c = 0; c < cp.numcomps; ++c)
        out.push_back({chunk, c, r.x0, y, r.w(), std::min(rows_per, r.y1 - y)});
    ti = tj;
  }
}
/* narrow one chunk piece by piece into the ring and send each piece on its way */
static int ring_upload_chunk(b2k_device_job* J, void* const* user, const uint32_t* strides, uint32_t chunk, cudaStream_t cs,
                             uint64_t& counter, const std::function<int()>& between_pieces)
{
  const b2k_coding& cp = J->cp;
  std::vector<StagePiece> pcs;
  chunk_pieces(J, chunk, pcs);
  for(const StagePiece& pc : pcs)
  {
    const uint32_t slot = (uint32_t)(counter % J->ring_slots);
    if(counter >= J->ring_slots)
      CUDA_TRY(cudaEventSynchronize(J->ring_ev[slot])); /* the slot's previous piece has left */
    uint16_t* sp = J->h_ring + (uint64_t)slot * J->ring_slot_elems;
    const int32_t* u = reinterpret_cast<const int32_t*>(user[pc.comp]) + (size_t)(pc.y0 - cp.y0) * strides[pc.comp] + (pc.x0 - cp.x0);
    const b2k_host_rect hr{u, sp, strides[pc.comp], pc.w, pc.w, pc.rows};
    b2k_host_convert(&hr, 1, false, cp.sgnd != 0, true);
    uint16_t* dev = J->img16.at(pc.comp, pc.x0, pc.y0);
    if(J->img16.pitch == pc.w)
      CUDA_TRY(cudaMemcpyAsync(dev, sp, (size_t)pc.w * pc.rows * 2, cudaMemcpyHostToDevice, cs));
    else
      CUDA_TRY(cudaMemcpy2DAsync(dev, (size_t)J->img16.pitch * 2, sp, (size_t)pc.w * 2, (size_t)pc.w * 2, pc.rows,
                                 cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaEventRecord(J->ring_ev[slot], cs));
    ++counter;
    if(between_pieces())
      return -1;
  }
  return 0;
}
/* bring every chunk's pixels down through the ring and widen them into the caller's planes; chunk k's pixels are
   ready on the device when chunk_ev[CEV(0, k)] fires */
static int ring_download_all(b2k_device_job* J, void* const* user, const uint32_t* strides, cudaStream_t cs)
{
  const b2k_coding& cp = J->cp;
  std::vector<StagePiece> pcs;
  for(uint32_t k = 0; k + 1 < J->chunk_tile.size(); ++k)
    chunk_pieces(J, k, pcs);
  const size_t N = pcs.size(), S = J->ring_slots;
  auto issue = [&](size_t p) -> int {
    const StagePiece& pc = pcs[p];
    const uint32_t slot = (uint32_t)(p % S);
    uint16_t* sp = J->h_ring + (uint64_t)slot * J->ring_slot_elems;
    if(p == 0 || pcs[p - 1].chunk != pc.chunk)
      CUDA_TRY(cudaStreamWaitEvent(cs, J->chunk_ev[CEV(0, pc.chunk)], 0));
    const uint16_t* dev = J->img16.at(pc.comp, pc.x0, pc.y0);
    if(J->img16.pitch == pc.w)
      CUDA_TRY(cudaMemcpyAsync(sp, dev, (size_t)pc.w * pc.rows * 2, cudaMemcpyDeviceToHost, cs));
    else
      CUDA_TRY(cudaMemcpy2DAsync(sp, (size_t)pc.w * 2, dev, (size_t)J->img16.pitch * 2, (size_t)pc.w * 2, pc.rows,
                                 cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(cudaEventRecord(J->ring_ev[slot], cs));
    return 0;
  };
  for(size_t p = 0; p < std::min(S, N); ++p)
    if(issue(p)) return -1;
  for(size_t p = 0; p < N; ++p)
  {
    const StagePiece& pc = pcs[p];
    const uint32_t slot = (uint32_t)(p % S);
    CUDA_TRY(cudaEventSynchronize(J->ring_ev[slot]));
    const uint16_t* sp = J->h_ring + (uint64_t)slot * J->ring_slot_elems;
    int32_t* u = reinterpret_cast<int32_t*>(user[pc.comp]) + (size_t)(pc.y0 - cp.y0) * strides[pc.comp] + (pc.x0 - cp.x0);
    const b2k_host_rect hr{sp, u, pc.w, strides[pc.comp], pc.w, pc.rows};
    b2k_host_convert(&hr, 1, true, cp.sgnd != 0);
    if(p + S < N)
      if(issue(p + S)) return -1;
  }
  return 0;
}

/* ---- stages ----------------------------------------------------------------------------------- */
static int enqueue_forward(b2k_device_job* J, cudaStream_t st, bool time_level1, size_t t0 = 0, size_t t1 = (size_t)-1,
                           bool use16 = false, cudaEvent_t l1_begin = nullptr, cudaEvent_t l1_end = nullptr)
{
  if(!l1_begin) l1_begin = J->ev[4];
  if(!l1_end) l1_end = J->ev[5];
  t1 = std::min(t1, J->tiles.size());
  bool first = true;
  for(size_t li = 0; li < J->fwd.size(); ++li)
  {
    const bool l16 = false;
    (void)use16;
    LevelLaunch& L = J->fwd[li];
    const uint32_t d0 = L.tile_first[t0], d1 = L.tile_first[t1];
    if(first && time_level1)
      CUDA_TRY(cudaEventRecord(l1_begin, st));
    if(d1 > d0 && L.point)
      b2k_launch_point_transform(L.d_descs + d0, (int)(d1 - d0), L.max_w, L.max_h, L.nc, J->cp.irreversible, true, st);
    else if(d1 > d0)
      b2k_launch_dwt_fwd(L.d_descs + d0, (int)(d1 - d0), L.max_jobs, L.nc, J->cp.irreversible, l16, st);
    if(first && time_level1)
    {
      CUDA_TRY(cudaEventRecord(l1_end, st));
      J->level1_alg_bytes = L.alg_bytes;
    }
    first = false;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int enqueue_inverse(b2k_device_job* J, cudaStream_t st, size_t t0 = 0, size_t t1 = (size_t)-1, bool use16 = false)
{
  t1 = std::min(t1, J->tiles.size());
  for(size_t li = 0; li < J->inv.size(); ++li)
  {
    const bool l16 = false;
    (void)use16;
    LevelLaunch& L = J->inv[li];
    const uint32_t d0 = L.tile_first[t0], d1 = L.tile_first[t1];
    if(d1 > d0 && L.point)
      b2k_launch_point_transform(L.d_descs + d0, (int)(d1 - d0), L.max_w, L.max_h, L.nc, J->cp.irreversible, false, st);
    else if(d1 > d0)
      b2k_launch_dwt_inv(L.d_descs + d0, (int)(d1 - d0), L.max_jobs, L.nc, J->cp.irreversible, l16, st);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

/* block coder over the coded blocks of selected tiles [t0, t1) */
static int enqueue_t1_blocks(b2k_device_job* J, cudaStream_t st, size_t t0, size_t t1)
{
  t1 = std::min(t1, J->tiles.size());
  const uint32_t b0 = J->coded_first[t0], b1 = J->coded_first[t1];
  if(b1 > b0)
    b2k_launch_ht_encode(J->d_enc_desc + b0, J->d_out + b0, J->d_scratch, b1 - b0, J->enc_limits, J->cp.irreversible, st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int enqueue_t1_encode(b2k_device_job* J, cudaStream_t st)
{
  const uint32_t n = (uint32_t)J->h_enc_desc.size();
  b2k_launch_ht_encode(J->d_enc_desc, J->d_out, J->d_scratch, n, J->enc_limits, J->cp.irreversible, st);
  b2k_launch_scan_lengths(J->d_out, J->d_offsets, n, st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

/* total size known -> (re)allocate the arena, compact */
static int finish_t1_encode(b2k_device_job* J, cudaStream_t st)
{
  const uint32_t n = (uint32_t)J->h_enc_desc.size();
  uint64_t total = 0;
  CUDA_TRY(cudaMemcpyAsync(&J->h_offsets[n], J->d_offsets + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  total = J->h_offsets[n];
  if(total + 64 > J->bytes_cap)
  {
    cudaFree(J->d_bytes);
    J->bytes_cap = total + total / 8 + 4096;
    CUDA_TRY(cudaMalloc(&J->d_bytes, J->bytes_cap));
  }
  J->bytes_used = total;
  b2k_launch_ht_gather(J->d_enc_desc, J->d_out, J->d_offsets, J->d_scratch, J->d_bytes, n, J->bytes_cap, st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int32_t b2k_job_forward(b2k_device_job* J, float* ms)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  cudaStream_t st = J->eng->stream;
  CUDA_TRY(cudaEventRecord(J->ev[0], st));
  if(enqueue_forward(J, st, true)) return -1;
  CUDA_TRY(cudaEventRecord(J->ev[1], st));
  CUDA_TRY(cudaEventSynchronize(J->ev[1]));
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, J->ev[0], J->ev[1]));
  if(ms) *ms = t;
  CUDA_TRY(cudaEventElapsedTime(&J->last_level1_ms, J->ev[4], J->ev[5]));
  return 0;
}

extern "C" int32_t b2k_job_inverse(b2k_device_job* J, float* ms)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  cudaStream_t st = J->eng->stream;
  CUDA_TRY(cudaEventRecord(J->ev[0], st));
  if(enqueue_inverse(J, st)) return -1;
  CUDA_TRY(cudaEventRecord(J->ev[1], st));
  CUDA_TRY(cudaEventSynchronize(J->ev[1]));
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, J->ev[0], J->ev[1]));
  if(ms) *ms = t;
  return 0;
}

extern "C" int32_t b2k_job_t1_encode(b2k_device_job* J, float* ms, uint64_t* total_bytes)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  cudaStream_t st = J->eng->stream;
  CUDA_TRY(cudaEventRecord(J->ev[0], st));
  if(enqueue_t1_encode(J, st)) return -1;
  if(finish_t1_encode(J, st)) return -1;
  CUDA_TRY(cudaEventRecord(J->ev[1], st));
  CUDA_TRY(cudaEventSynchronize(J->ev[1]));
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, J->ev[0], J->ev[1]));
  if(ms) *ms = t;
  if(total_bytes) *total_bytes = J->bytes_used;
  return 0;
}

/* decode descriptors from (length, offset, numbps) per coded block, for coded blocks [k0, k1) */
static int prepare_decode(b2k_device_job* J, const b2k_block* blocks, uint64_t num_blocks, cudaStream_t st, size_t k0 = 0,
                          size_t k1 = (size_t)-1)
{
  const size_t n = J->h_enc_desc.size();
  k1 = std::min(k1, n);
  if(num_blocks != J->blocks.size())
  {
    g_err = "block count does not match this coding's enumeration";
    return -1;
  }
  for(size_t k = k0; k < k1; ++k)
  {
    const b2k_block& b = blocks[J->coded_index[k]];
    HtBlockDesc d = J->h_enc_desc[k];
    d.length = b.length;
    d.slot_off = b.offset;
    if(b.numpasses > 3)
    {
      g_err = "an HT code block with more than 3 coding passes";
      return 1;
    }
    const int nb = b.length ? b.numbps : 0;
    d.mmsbs = (uint8_t)std::max(0, (int)d.kmax - nb);
    /* ojph_block_decoder32.cpp L752-758, L790-803: no refinement bytes, or a cleanup pass already at
       bit-plane 1, leave nothing to refine */
    d.passes = (b.length && b.numpasses > 1 && b.length2 > 0 && d.mmsbs < 29) ? b.numpasses : 1;
    d.length2 = d.passes > 1 ? b.length2 : 0;
    if(d.passes > 1)
      J->dec_has_refinement = true;
    d.quant = J->dec_quant[k]; /* stepsize / 2^(31-Kmax), PostDecodeFiltersOJPH.h L103 */
    J->h_dec_desc[k] = d;
  }
  if(k1 > k0)
    CUDA_TRY(cudaMemcpyAsync(J->d_dec_desc + k0, J->h_dec_desc + k0, (k1 - k0) * sizeof(HtBlockDesc), cudaMemcpyHostToDevice, st));
  return 0;
}

static int enqueue_t1_decode_own(b2k_device_job* J, cudaStream_t st)
{
  const uint32_t n = (uint32_t)J->h_enc_desc.size();
  b2k_launch_build_dec_desc(J->d_enc_desc, J->d_out, J->d_offsets, J->d_dec_quant, J->d_dec_desc, n, st);
  b2k_launch_ht_decode(J->d_dec_desc, J->d_bytes, J->d_recs, J->d_dec_status, n, J->max_cblk_w, J->d_err, J->cp.irreversible, 0, st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int32_t b2k_job_t1_decode(b2k_device_job* J, float* ms)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  cudaStream_t st = J->eng->stream;
  CUDA_TRY(cudaMemsetAsync(J->d_err, 0, sizeof(int), st));
  CUDA_TRY(cudaEventRecord(J->ev[0], st));
  if(enqueue_t1_decode_own(J, st)) return -1;
  CUDA_TRY(cudaEventRecord(J->ev[1], st));
  CUDA_TRY(cudaEventSynchronize(J->ev[1]));
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, J->ev[0], J->ev[1]));
  if(ms) *ms = t;
  int herr = 0;
  CUDA_TRY(cudaMemcpy(&herr, J->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if(herr)
  {
    g_err = "HT decoder rejected " + std::to_string(herr) + " block(s)";
    return -2;
  }
  return 0;
}

/* stage hook: block-decode a caller-supplied block table (what the host's T2 parse produced, or a foreign
   stream's blocks with SigProp / MagRef passes) into the job's coefficient planes */
extern "C" int32_t b2k_job_t1_decode_blocks(b2k_device_job* J, const b2k_block* blocks, uint64_t num_blocks,
                                            const uint8_t* bytes, uint64_t num_bytes, float* ms)
{
  if(!J || !blocks || (!bytes && num_bytes)) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  cudaStream_t st = J->eng->stream;
  if(num_bytes + 64 > J->bytes_cap)
  {
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(J->d_bytes);
    J->bytes_cap = num_bytes + 4096;
    CUDA_TRY(cudaMalloc(&J->d_bytes, J->bytes_cap));
  }
  J->dec_has_refinement = false;
  if(int prc = prepare_decode(J, blocks, num_blocks, st)) return prc;
  const uint32_t n = (uint32_t)J->h_enc_desc.size();
  for(uint32_t k = 0; k < n; ++k)
    if((uint64_t)J->h_dec_desc[k].slot_off + J->h_dec_desc[k].length + J->h_dec_desc[k].length2 > num_bytes)
    {
      g_err = "block offsets exceed the byte arena";
      return -1;
    }
  CUDA_TRY(cudaMemsetAsync(J->d_err, 0, sizeof(int), st));
  if(num_bytes)
    CUDA_TRY(cudaMemcpyAsync(J->d_bytes, bytes, num_bytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(J->ev[0], st));
  b2k_launch_ht_decode(J->d_dec_desc, J->d_bytes, J->d_recs, J->d_dec_status, n, J->max_cblk_w, J->d_err, J->cp.irreversible,
                       J->dec_has_refinement, st);
  if(J->dec_has_refinement)
    b2k_launch_ht_decode_refine(J->d_dec_desc, J->d_bytes, J->d_dec_status, n, (J->cp.cblk_sty & 0x08) != 0, st);
  CUDA_TRY(cudaEventRecord(J->ev[1], st));
  CUDA_TRY(cudaEventSynchronize(J->ev[1]));
  CUDA_TRY(cudaGetLastError());
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, J->ev[0], J->ev[1]));
  if(ms) *ms = t;
  int herr = 0;
  CUDA_TRY(cudaMemcpy(&herr, J->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if(herr)
  {
    g_err = "HT decoder rejected " + std::to_string(herr) + " block(s)";
    return -2;
  }
  return 0;
}

/* One device-resident round trip, enqueued back to back with a single synchronisation at the end:
   forward (DC shift + MCT + DWT) -> block encode -> scan + compact -> block decode -> inverse.
   stage_ms[4] (optional) = forward, encode (incl. scan/compact), decode, inverse from CUDA events. */
extern "C" int32_t b2k_job_roundtrip(b2k_device_job* J, float* ms_total, float* stage_ms, uint64_t* total_bytes)
{
  if(!J) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  cudaStream_t st = J->eng->stream;
  const uint32_t n = (uint32_t)J->h_enc_desc.size();
  if(J->bytes_cap == 0)
  { /* first use: size the arena (one synchronising pass) */
    float t;
    uint64_t b;
    if(int rc = b2k_job_forward(J, &t)) return rc;
    if(int rc = b2k_job_t1_encode(J, &t, &b)) return rc;
  }
  CUDA_TRY(cudaMemsetAsync(J->d_err, 0, sizeof(int), st));
  CUDA_TRY(cudaEventRecord(J->ev[0], st));
  if(enqueue_forward(J, st, true)) return -1;
  CUDA_TRY(cudaEventRecord(J->ev[1], st));
  if(enqueue_t1_encode(J, st)) return -1;
  b2k_launch_ht_gather(J->d_enc_desc, J->d_out, J->d_offsets, J->d_scratch, J->d_bytes, n, J->bytes_cap, st);
  CUDA_TRY(cudaMemcpyAsync(&J->h_offsets[n], J->d_offsets + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(J->ev[2], st));
  if(enqueue_t1_decode_own(J, st)) return -1;
  CUDA_TRY(cudaEventRecord(J->ev[3], st));
  if(enqueue_inverse(J, st)) return -1;
  CUDA_TRY(cudaEventRecord(J->ev[6], st));
  CUDA_TRY(cudaEventSynchronize(J->ev[6]));
  CUDA_TRY(cudaGetLastError());
  if(J->h_offsets[n] > J->bytes_cap)
  { /* arena estimate too small (content changed): grow and let the caller repeat */
    cudaFree(J->d_bytes);
    J->bytes_cap = J->h_offsets[n] + J->h_offsets[n] / 8 + 4096;
    CUDA_TRY(cudaMalloc(&J->d_bytes, J->bytes_cap));
    g_err = "coded size grew past the arena: arena resized, call again";
    return 2;
  }
  J->bytes_used = J->h_offsets[n];
  float t[5] = {0, 0, 0, 0, 0};
  cudaEventElapsedTime(&t[0], J->ev[0], J->ev[1]);
  cudaEventElapsedTime(&t[1], J->ev[1], J->ev[2]);
  cudaEventElapsedTime(&t[2], J->ev[2], J->ev[3]);
  cudaEventElapsedTime(&t[3], J->ev[3], J->ev[6]);
  cudaEventElapsedTime(&t[4], J->ev[0], J->ev[6]);
  cudaEventElapsedTime(&J->last_level1_ms, J->ev[4], J->ev[5]);
  if(ms_total) *ms_total = t[4];
  if(stage_ms)
    for(int i = 0; i < 4; ++i)
      stage_ms[i] = t[i];
  if(total_bytes) *total_bytes = J->bytes_used;
  int herr = 0;
  CUDA_TRY(cudaMemcpy(&herr, J->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if(herr)
  {
    g_err = "HT decoder rejected " + std::to_string(herr) + " block(s)";
    return -2;
  }
  return 0;
}

/* n device-resident round trips queued back to back on the stream, ONE synchronisation after the last: what a
   benchmark step loop should cost when the host is not in the way (several ranks on one box).  Times come from
   events recorded per step: ms_total = first step's start to last step's end; stage_ms[4] and level1_ms are sums
   over the steps.  Returns 2 once if the coded size outgrew the arena (it has been resized: call again). */
extern "C" int32_t b2k_job_roundtrip_n(b2k_device_job* J, uint32_t steps, float* ms_total, float* stage_ms, float* level1_ms,
                                       uint64_t* total_bytes)
{
  if(!J || !steps) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  cudaStream_t st = J->eng->stream;
  const uint32_t n = (uint32_t)J->h_enc_desc.size();
  if(J->bytes_cap == 0)
  { /* first use: size the arena (one synchronising pass) */
    float t;
    uint64_t b;
    if(int rc = b2k_job_forward(J, &t)) return rc;
    if(int rc = b2k_job_t1_encode(J, &t, &b)) return rc;
  }
  const size_t per = 7; /* start, l1 begin, l1 end, after forward, after encode, after decode, end */
  while(J->q_ev.size() < per * steps)
  {
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreate(&ev));
    J->q_ev.push_back(ev);
  }
  CUDA_TRY(cudaMemsetAsync(J->d_err, 0, sizeof(int), st));
  for(uint32_t s = 0; s < steps; ++s)
  {
    cudaEvent_t* e = J->q_ev.data() + per * s;
    CUDA_TRY(cudaEventRecord(e[0], st));
    if(enqueue_forward(J, st, true, 0, (size_t)-1, false, e[1], e[2])) return -1;
    CUDA_TRY(cudaEventRecord(e[3], st));
    if(enqueue_t1_encode(J, st)) return -1;
    b2k_launch_ht_gather(J->d_enc_desc, J->d_out, J->d_offsets, J->d_scratch, J->d_bytes, n, J->bytes_cap, st);
    if(s + 1 == steps)
      CUDA_TRY(cudaMemcpyAsync(&J->h_offsets[n], J->d_offsets + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(e[4], st));
    if(enqueue_t1_decode_own(J, st)) return -1;
    CUDA_TRY(cudaEventRecord(e[5], st));
    if(enqueue_inverse(J, st)) return -1;
    CUDA_TRY(cudaEventRecord(e[6], st));
  }
  CUDA_TRY(cudaEventSynchronize(J->q_ev[per * (steps - 1) + 6]));
  CUDA_TRY(cudaGetLastError());
  if(J->h_offsets[n] > J->bytes_cap)
  {
    cudaFree(J->d_bytes);
    J->bytes_cap = J->h_offsets[n] + J->h_offsets[n] / 8 + 4096;
    CUDA_TRY(cudaMalloc(&J->d_bytes, J->bytes_cap));
    g_err = "coded size grew past the arena: arena resized, call again";
    return 2;
  }
  J->bytes_used = J->h_offsets[n];
  float sums[4] = {0, 0, 0, 0}, l1 = 0, tot = 0;
  for(uint32_t s = 0; s < steps; ++s)
  {
    cudaEvent_t* e = J->q_ev.data() + per * s;
    float t = 0;
    cudaEventElapsedTime(&t, e[0], e[3]); sums[0] += t;
    cudaEventElapsedTime(&t, e[3], e[4]); sums[1] += t;
    cudaEventElapsedTime(&t, e[4], e[5]); sums[2] += t;
    cudaEventElapsedTime(&t, e[5], e[6]); sums[3] += t;
    cudaEventElapsedTime(&t, e[1], e[2]); l1 += t;
  }
  cudaEventElapsedTime(&tot, J->q_ev[0], J->q_ev[per * (steps - 1) + 6]);
  J->last_level1_ms = l1 / steps;
  if(ms_total) *ms_total = tot;
  if(stage_ms)
    for(int i = 0; i < 4; ++i)
      stage_ms[i] = sums[i];
  if(level1_ms) *level1_ms = l1;
  if(total_bytes) *total_bytes = J->bytes_used;
  int herr = 0;
  CUDA_TRY(cudaMemcpy(&herr, J->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if(herr)
  {
    g_err = "HT decoder rejected " + std::to_string(herr) + " block(s)";
    return -2;
  }
  return 0;
}

/* The same n round trips with the BLOCK-CODER stage software-pipelined over tile-independent block ranges: the forward
   transform of the whole image runs alone on the main stream (its kernels are HBM-bound and are timed as such), then the
   coded blocks are cut into `chunks` ranges (0: 2) and each range goes encode -> scan -> compact -> parse (phase A) -> MagSgn
   (phase B) on one of `streams` side streams, so that the latency-bound kernels of one range (phase A is a serial chain
   per block, the scan a single CTA) run under the issue-bound kernels of its neighbours; the inverse transform again
   runs alone after all ranges have joined.  Ranges are independent except for the byte offsets of the compacted arena,
   which chain from one range's scan to the next (event per range).  stage_ms[3] = forward, block coder (encode + decode
   together), inverse.  Same results as b2k_job_roundtrip_n, byte for byte. */
extern "C" int32_t b2k_job_roundtrip_pipelined_n(b2k_device_job* J, uint32_t steps, uint32_t chunks, uint32_t streams, float* ms_total,
                                                 float* stage_ms, float* level1_ms, uint64_t* total_bytes)
{
  if(!J || !steps) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  cudaStream_t st = J->eng->stream;
  const uint32_t n = (uint32_t)J->h_enc_desc.size();
  if(J->bytes_cap == 0)
  { /* first use: size the arena (one synchronising pass) */
    float t;
    uint64_t b;
    if(int rc = b2k_job_forward(J, &t)) return rc;
    if(int rc = b2k_job_t1_encode(J, &t, &b)) return rc;
  }
  /* measured on config 2 (tools/stage_times.py): 2 ranges on 2 streams 2.62 ms against 2.82 ms back to
     back; more ranges per stream lose (phase A's serial chain is a ~0.35 ms floor per launch, the persistent encoder grid
     fills every SM's shared memory), and so does a dedicated encoder stream with a capped grid */
  chunks = std::max(1u, std::min(chunks ? chunks : 2u, 64u));
  streams = std::max(1u, std::min(streams ? streams : 2u, 8u));
  /* ranges start on multiples of 128 blocks: whole CTAs of both decode kernels, whole record-interleave groups */
  const uint32_t per_chunk = std::max(128u, (((n + chunks - 1) / chunks) + 127u) & ~127u);
  const uint32_t nch = n ? (n + per_chunk - 1) / per_chunk : 0;
  while(J->p_streams.size() < streams)
  {
    cudaStream_t ps;
    CUDA_TRY(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    J->p_streams.push_back(ps);
  }
  while(J->p_ev.size() < 2 * (size_t)nch)
  {
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    J->p_ev.push_back(ev);
  }
  const size_t per = 5; /* start, l1 begin, l1 end, after forward, after the block coder; the next step's start ends the inverse */
  while(J->q_ev.size() < per * steps + 1)
  {
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreate(&ev));
    J->q_ev.push_back(ev);
  }
  CUDA_TRY(cudaMemsetAsync(J->d_err, 0, sizeof(int), st));
  for(uint32_t s = 0; s < steps; ++s)
  {
    cudaEvent_t* e = J->q_ev.data() + per * s;
    CUDA_TRY(cudaEventRecord(e[0], st));
    if(enqueue_forward(J, st, true, 0, (size_t)-1, false, e[1], e[2])) return -1;
    CUDA_TRY(cudaEventRecord(e[3], st));
    for(uint32_t k = 0; k < streams && k < nch; ++k)
      CUDA_TRY(cudaStreamWaitEvent(J->p_streams[k], e[3], 0));
    for(uint32_t c = 0; c < nch; ++c)
    {
      cudaStream_t ps = J->p_streams[c % streams];
      const uint32_t b0 = c * per_chunk, b1 = std::min(n, b0 + per_chunk), nb = b1 - b0;
      b2k_launch_ht_encode(J->d_enc_desc + b0, J->d_out + b0, J->d_scratch, nb, J->enc_limits, J->cp.irreversible, ps);
      if(c > 0)
        CUDA_TRY(cudaStreamWaitEvent(ps, J->p_ev[2 * (c - 1)], 0)); /* the previous range's end offset */
      b2k_launch_scan_lengths(J->d_out + b0, J->d_offsets + b0, nb, ps);
      CUDA_TRY(cudaEventRecord(J->p_ev[2 * c], ps));
      b2k_launch_ht_gather(J->d_enc_desc + b0, J->d_out + b0, J->d_offsets + b0, J->d_scratch, J->d_bytes, nb, J->bytes_cap, ps);
      b2k_launch_build_dec_desc(J->d_enc_desc + b0, J->d_out + b0, J->d_offsets + b0, J->d_dec_quant + b0, J->d_dec_desc + b0, nb, ps);
      b2k_launch_ht_decode(J->d_dec_desc + b0, J->d_bytes, J->d_recs, J->d_dec_status + b0, nb, J->max_cblk_w, J->d_err,
                           J->cp.irreversible, 0, ps);
      CUDA_TRY(cudaEventRecord(J->p_ev[2 * c + 1], ps));
    }
    for(uint32_t c = 0; c < nch; ++c)
      CUDA_TRY(cudaStreamWaitEvent(st, J->p_ev[2 * c + 1], 0));
    if(s + 1 == steps)
      CUDA_TRY(cudaMemcpyAsync(&J->h_offsets[n], J->d_offsets + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(e[4], st));
    if(enqueue_inverse(J, st)) return -1;
  }
  cudaEvent_t last = J->q_ev[per * steps];
  CUDA_TRY(cudaEventRecord(last, st));
  CUDA_TRY(cudaEventSynchronize(last));
  CUDA_TRY(cudaGetLastError());
  if(J->h_offsets[n] > J->bytes_cap)
  {
    cudaFree(J->d_bytes);
    J->bytes_cap = J->h_offsets[n] + J->h_offsets[n] / 8 + 4096;
    CUDA_TRY(cudaMalloc(&J->d_bytes, J->bytes_cap));
    g_err = "coded size grew past the arena: arena resized, call again";
    return 2;
  }
  J->bytes_used = J->h_offsets[n];
  float sums[3] = {0, 0, 0}, l1 = 0, tot = 0;
  for(uint32_t s = 0; s < steps; ++s)
  {
    cudaEvent_t* e = J->q_ev.data() + per * s;
    float t = 0;
    cudaEventElapsedTime(&t, e[0], e[3]); sums[0] += t;
    cudaEventElapsedTime(&t, e[3], e[4]); sums[1] += t;
    cudaEventElapsedTime(&t, e[4], e[per]); sums[2] += t; /* e[per] = the next step's start, or `last` */
    cudaEventElapsedTime(&t, e[1], e[2]); l1 += t;
  }
  cudaEventElapsedTime(&tot, J->q_ev[0], last);
  J->last_level1_ms = l1 / steps;
  if(ms_total) *ms_total = tot;
  if(stage_ms)
    for(int i = 0; i < 3; ++i)
      stage_ms[i] = sums[i];
  if(level1_ms) *level1_ms = l1;
  if(total_bytes) *total_bytes = J->bytes_used;
  int herr = 0;
  CUDA_TRY(cudaMemcpy(&herr, J->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if(herr)
  {
    g_err = "HT decoder rejected " + std::to_string(herr) + " block(s)";
    return -2;
  }
  return 0;
}

extern "C" int32_t b2k_job_last_kernel_stats(const b2k_device_job* J, int which, float* ms, uint64_t* alg_bytes)
{
  if(!J) return -1;
  if(which == 0)
  {
    if(ms) *ms = J->last_level1_ms;
    if(alg_bytes) *alg_bytes = J->level1_alg_bytes;
    return 0;
  }
  return 1;
}

/* ---- results ---------------------------------------------------------------------------------- */
/* pinned result arenas are recycled: cudaHostAlloc of ~150 MB costs more than the whole encode */
struct PinnedPool
{
  std::mutex mu;
  std::vector<std::pair<uint8_t*, uint64_t>> free_list;
  std::vector<std::pair<uint8_t*, uint64_t>> live;
  std::vector<uint8_t*> heap; /* arenas that had to come from malloc (no CUDA device: b2k_result_merge on a writer-only host) */
};
static PinnedPool g_pool;
static uint8_t* pool_get(uint64_t bytes)
{
  std::lock_guard<std::mutex> lock(g_pool.mu);
  for(size_t i = 0; i < g_pool.free_list.size(); ++i)
    if(g_pool.free_list[i].second >= bytes)
    {
      auto e = g_pool.free_list[i];
      g_pool.free_list.erase(g_pool.free_list.begin() + i);
      g_pool.live.push_back(e);
      return e.first;
    }
  uint8_t* p = nullptr;
  const uint64_t cap = bytes + bytes / 4 + 4096;
  if(cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess)
    return nullptr;
  g_pool.live.push_back({p, cap});
  return p;
}
static void pool_put(uint8_t* p)
{
  std::lock_guard<std::mutex> lock(g_pool.mu);
  for(size_t i = 0; i < g_pool.heap.size(); ++i)
    if(g_pool.heap[i] == p)
    {
      g_pool.heap.erase(g_pool.heap.begin() + i);
      free(p);
      return;
    }
  for(size_t i = 0; i < g_pool.live.size(); ++i)
    if(g_pool.live[i].first == p)
    {
      g_pool.free_list.push_back(g_pool.live[i]);
      g_pool.live.erase(g_pool.live.begin() + i);
      while(g_pool.free_list.size() > 16) /* streams keep depth-many results alive per direction; re-pinning 150 MB costs tens of ms */
      {
        cudaFreeHost(g_pool.free_list.front().first);
        g_pool.free_list.erase(g_pool.free_list.begin());
      }
      return;
    }
  cudaFreeHost(p);
}

/* the part of a result that does not depend on the device: the block table in enumeration order */
static b2k_result* result_shell(b2k_device_job* J)
{
  b2k_result* R = new b2k_result();
  memset(R, 0, sizeof(*R));
  R->num_blocks = J->blocks.size();
  R->blocks = (b2k_block*)malloc(sizeof(b2k_block) * std::max<size_t>(1, J->blocks.size()));
  memcpy(R->blocks, J->blocks.data(), sizeof(b2k_block) * J->blocks.size());
  R->num_tiles = (uint32_t)J->tiles.size();
  return R;
}

/* host_bytes: the caller already brought the byte arena home (chunk by chunk); meta_on_host: also the per-block
   lengths / offsets (h_out, h_offsets) are on their way on a stream that `st` waits for */
static int fetch_result(b2k_device_job* J, cudaStream_t st, b2k_result** out, uint8_t* host_bytes = nullptr,
                        b2k_result* shell = nullptr, bool meta_on_host = false)
{
  const uint32_t n = (uint32_t)J->h_enc_desc.size();
  b2k_result* R = shell ? shell : result_shell(J);
  R->num_bytes = J->bytes_used;
  R->bytes = host_bytes ? host_bytes : pool_get(std::max<uint64_t>(64, J->bytes_used));
  if(!R->bytes)
  {
    g_err = "cudaHostAlloc(result bytes) failed";
    free(R->blocks);
    delete R;
    return -1;
  }
  struct ResultGuard /* a failing CUDA call below must not leak the result and its pinned arena */
  {
    b2k_result* r;
    ~ResultGuard() { if(r) b2k_result_free(r); }
  } guard{R};
  if(!host_bytes)
    CUDA_TRY(cudaMemcpyAsync(R->bytes, J->d_bytes, J->bytes_used, cudaMemcpyDeviceToHost, st));
  if(!meta_on_host)
  {
    CUDA_TRY(cudaMemcpyAsync(J->h_out, J->d_out, n * sizeof(HtBlockOut), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(J->h_offsets, J->d_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  int bad = 0;
  for(uint32_t k = 0; k < n; ++k)
  {
    b2k_block& b = R->blocks[J->coded_index[k]];
    if(J->h_out[k].total == 0xFFFFFFFFu)
    {
      bad++;
      continue;
    }
    b.length = J->h_out[k].total;
    b.offset = J->h_offsets[k];
    b.numbps = 1;    /* CoderOJPH.cpp L203-206 */
    b.numpasses = 1;
  }
  if(bad)
  {
    g_err = std::to_string(bad) + " code block(s) overflowed the coder's buffers";
    return -2;
  }
  guard.r = nullptr;
  *out = R;
  return 0;
}

extern "C" int32_t b2k_job_fetch_result(b2k_device_job* J, b2k_result** out)
{
  if(!J || !out) return -1;
  CUDA_TRY(cudaSetDevice(J->eng->device));
  return fetch_result(J, J->eng->stream, out);
}

extern "C" void b2k_result_free(b2k_result* r)
{
  if(!r) return;
  free(r->blocks);
  if(r->bytes) pool_put(r->bytes);
  delete r;
}

/* Merge the results of ranks that each coded the tiles t with t % nshards == rank (tile_mod / tile_rem of b2k_encode)
   into one result in full enumeration order -- what the writer rank needs for b2k_codestream_write after gathering the
   shards' block tables and byte arenas.  The shards are not modified; free the merged result with
   b2k_result_free. */
extern "C" int32_t b2k_result_merge(const b2k_coding* cp, const b2k_result* const* shards, uint32_t nshards, b2k_result** out)
{
  if(!cp || !shards || !nshards || !out)
    return -1;
  if(const char* why = unsupported_reason(*cp))
  {
    g_err = why;
    return -1;
  }
  const TileGrid g = tile_grid(*cp);
  const uint32_t ntiles = g.nx * g.ny;
  const std::vector<BandQuant> q = band_quant(*cp);
  std::vector<b2k_block> all;
  for(uint32_t t = 0; t < ntiles; ++t)
    enumerate_tile_blocks(*cp, t, tile_rect(*cp, g, t), q, all);
  uint64_t total_bytes = 0;
  std::vector<uint64_t> base(nshards, 0);
  for(uint32_t s = 0; s < nshards; ++s)
  {
    if(!shards[s])
    {
      g_err = "missing shard";
      return -1;
    }
    base[s] = total_bytes;
    total_bytes += shards[s]->num_bytes;
  }
  b2k_result* R = new b2k_result();
  memset(R, 0, sizeof(*R));
  R->num_blocks = all.size();
  R->num_tiles = ntiles;
  R->num_bytes = total_bytes;
  R->blocks = (b2k_block*)malloc(sizeof(b2k_block) * std::max<size_t>(1, all.size()));
  /* recycled pinned arena (pinning a gigabyte costs hundreds of milliseconds); plain memory on a writer-only host */
  uint8_t* arena = pool_get(std::max<uint64_t>(64, total_bytes));
  if(!arena)
  {
    (void)cudaGetLastError();
    arena = (uint8_t*)malloc(std::max<uint64_t>(64, total_bytes));
    std::lock_guard<std::mutex> lock(g_pool.mu);
    g_pool.heap.push_back(arena);
  }
  R->bytes = arena;
  if(!R->blocks || !arena)
  {
    b2k_result_free(R);
    g_err = "out of memory";
    return -1;
  }
  std::vector<uint64_t> next(nshards, 0);
  for(size_t i = 0; i < all.size(); ++i)
  {
    const uint32_t s = all[i].tile % nshards;
    const b2k_result* S = shards[s];
    if(next[s] >= S->num_blocks)
    {
      b2k_result_free(R);
      g_err = "a shard holds fewer blocks than its tiles have";
      return -1;
    }
    const b2k_block& b = S->blocks[next[s]++];
    if(b.tile != all[i].tile || b.comp != all[i].comp || b.resno != all[i].resno || b.band_index != all[i].band_index ||
       b.precno != all[i].precno || b.cblkno != all[i].cblkno)
    {
      b2k_result_free(R);
      g_err = "a shard's block table is not the enumeration of the tiles t % nshards == shard";
      return -1;
    }
    R->blocks[i] = b;
    R->blocks[i].offset = b.offset + base[s];
  }
  for(uint32_t s = 0; s < nshards; ++s)
  {
    if(next[s] != shards[s]->num_blocks)
    {
      b2k_result_free(R);
      g_err = "a shard holds more blocks than its tiles have";
      return -1;
    }
  }
  { /* the arenas, 8 MB pieces on the host pool */
    struct Piece { uint8_t* dst; const uint8_t* src; size_t n; };
    std::vector<Piece> pieces;
    const size_t step = (size_t)8 << 20;
    for(uint32_t s = 0; s < nshards; ++s)
      for(size_t o = 0; o < shards[s]->num_bytes; o += step)
        pieces.push_back({arena + base[s] + o, shards[s]->bytes + o, std::min<size_t>(step, shards[s]->num_bytes - o)});
    b2k_host_parallel(pieces.size(), [&](size_t i) { memcpy(pieces[i].dst, pieces[i].src, pieces[i].n); });
  }
  *out = R;
  return 0;
}

/* ---- one-call host paths ---------------------------------------------------------------------- */
static bool dbg_timing()
{
  static const bool v = getenv("B2K_DEBUG_TIMING") != nullptr;
  return v;
}
#define DBG_T(label)                                                                                              \
  do                                                                                                              \
  {                                                                                                               \
    if(dbg_timing())                                                                                              \
      fprintf(stderr, "[b2k] %-28s %8.3f ms\n", label,                                                            \
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());       \
  } while(0)

static b2k_device_job* cached_job(b2k_engine* e, const b2k_coding* cp, uint32_t mod, uint32_t rem, int* rc)
{
  b2k_device_job*& J = e->cached;
  cudaSetDevice(e->device);
  if(J && (memcmp(&J->cp, cp, sizeof(b2k_coding)) != 0 || J->tile_mod != mod || J->tile_rem != rem))
  {
    b2k_job_destroy(J);
    J = nullptr;
  }
  if(!J)
    *rc = b2k_job_create(e, cp, mod, rem, &J);
  else
    *rc = 0;
  return J;
}

static int32_t encode_common(b2k_engine* e, const b2k_coding* cp, void* const* planes, const uint32_t* strides,
                             uint32_t mod, uint32_t rem, b2k_result** out, bool u16, bool interleaved = false)
{
  if(!e || !cp || !planes || !strides || !out)
    return -1;
  std::lock_guard<std::mutex> lock(e->mu);
  int rc = 0;
  b2k_device_job* J = cached_job(e, cp, mod, rem, &rc);
  if(rc)
    return rc;
  CUDA_TRY(cudaSetDevice(e->device));
  void* const* user_planes = planes;
  const uint32_t* user_strides = strides;
  void* stage_planes[4];
  uint32_t stage_strides[4];
  /* several ranks on one host share its DRAM and CPU quota: measured packing loses there,
     so the automatic policy only considers it for a process that has the host to itself */
  const bool tuned = !u16 && host_pack_eligible(J) && g_pack_policy.load() < 0 && b2k_host_local_peers() == 1;
  const bool pack = !u16 && host_pack_eligible(J) && (tuned ? J->tune_enc.next_mode() : g_pack_policy.load() > 0);
  const auto wall0 = std::chrono::steady_clock::now();
  if(!u16)
    g_last_pack[0].store(pack ? 1 : 0);
  const bool ring = pack && ring_geom(false).slot_mb > 0;
  uint64_t ring_counter = 0;
  if(pack)
  {
    if(ring ? ensure_ring(J, false) : ensure_stage16(J)) return -1;
    if(!ring)
    {
      stage16_views(J, stage_planes, stage_strides);
      planes = stage_planes;
      strides = stage_strides;
    }
    u16 = true;
    b2k_host_session(true);
  }
  struct SessionEnd
  {
    bool on;
    ~SessionEnd() { if(on) b2k_host_session(false); }
  } session_end{pack};
  if(u16)
    if(int urc = interleaved ? ensure_ileave(J) : ensure_u16(J))
      return urc;
  cudaStream_t st = e->stream;
  /* software pipeline over tile chunks: chunk k+1 crosses PCIe on the copy stream while chunk k
     is transformed and block-coded on the compute stream */
  cudaStream_t cs = e->copy_stream;
  CUDA_TRY(cudaEventRecord(J->ev[0], st));
  CUDA_TRY(cudaStreamWaitEvent(cs, J->ev[0], 0));
  const size_t nchunks = J->chunk_tile.size() - 1;
  /* the arena size of the previous call is the estimate: scan + compact + return every chunk's bytes while later
     chunks are still arriving (the D2H direction of PCIe is otherwise idle) */
  const bool streamed = J->bytes_cap > 0 && nchunks > 1;
  uint8_t* hb = nullptr;
  struct ArenaGuard /* the pinned arena goes back to the pool on every early return until a result owns it */
  {
    uint8_t*& p;
    ~ArenaGuard() { if(p) pool_put(p); }
  } arena_guard{hb};
  if(streamed)
  {
    hb = pool_get(J->bytes_cap);
    if(!hb)
    {
      g_err = "cudaHostAlloc(result bytes) failed";
      return -1;
    }
  }
  cudaStream_t ds = e->h2d_stream; /* third stream: device-to-host here */
  size_t next_out = 0, enqueued = 0;
  uint64_t total = 0;
  bool overflow = false;
  /* send finished chunks' bytes home; non-blocking while the host still has chunks to feed */
  auto return_chunks = [&](bool block) -> int {
    while(next_out < enqueued)
    {
      const size_t k = next_out;
      if(block)
        CUDA_TRY(cudaEventSynchronize(J->chunk_ev[CEV(2, k)]));
      else if(cudaEventQuery(J->chunk_ev[CEV(2, k)]) != cudaSuccess)
        break;
      const uint32_t b0 = J->coded_first[J->chunk_tile[k]], b1 = J->coded_first[J->chunk_tile[k + 1]];
      const uint64_t lo = k == 0 ? 0 : J->h_offsets[b0], hi = J->h_offsets[b1];
      total = hi;
      if(hi > J->bytes_cap)
        overflow = true;
      else if(hi > lo)
      {
        CUDA_TRY(cudaStreamWaitEvent(ds, J->chunk_ev[CEV(2, k)], 0));
        CUDA_TRY(cudaMemcpyAsync(hb + lo, J->d_bytes + lo, hi - lo, cudaMemcpyDeviceToHost, ds));
      }
      if(b1 > b0)
      { /* per-block lengths and offsets of the chunk ride along (h_offsets[b1] is already here) */
        CUDA_TRY(cudaStreamWaitEvent(ds, J->chunk_ev[CEV(2, k)], 0));
        CUDA_TRY(cudaMemcpyAsync(J->h_out + b0, J->d_out + b0, (size_t)(b1 - b0) * sizeof(HtBlockOut), cudaMemcpyDeviceToHost, ds));
        CUDA_TRY(cudaMemcpyAsync(J->h_offsets + b0, J->d_offsets + b0, (size_t)(b1 - b0) * sizeof(uint64_t),
                                 cudaMemcpyDeviceToHost, ds));
      }
      ++next_out;
    }
    (void)cudaGetLastError(); /* cudaEventQuery's cudaErrorNotReady is not an error */
    return 0;
  };
  for(size_t k = 0; k < nchunks; ++k)
  {
    const size_t t0 = J->chunk_tile[k], t1 = J->chunk_tile[k + 1];
    if(ring)
    {
      if(ring_upload_chunk(J, user_planes, user_strides, (uint32_t)k, cs, ring_counter, [&] { return return_chunks(false); })) return -1;
    }
    else
    {
      if(pack)
        host_convert_chunk(J, user_planes, user_strides, false, t0, t1); /* overlaps chunk k-1's H2D */
      if(interleaved ? upload_interleaved(J, static_cast<const uint16_t*>(planes[0]), strides[0], cs, t0, t1)
         : u16       ? copy_planes16(J, planes, strides, true, cs, t0, t1)
                     : copy_planes(J, J->img, planes, strides, true, cs, t0, t1))
        return -1;
    }
    CUDA_TRY(cudaEventRecord(J->chunk_ev[CEV(0, k)], cs));
    CUDA_TRY(cudaStreamWaitEvent(st, J->chunk_ev[CEV(0, k)], 0));
    if(k == nchunks - 1)
      CUDA_TRY(cudaEventRecord(J->ev[1], st)); /* all planes on the device */
    if(u16 && (interleaved ? split_interleaved(J, st, t0, t1) : convert_planes16(J, true, st, t0, t1))) return -1;
    if(enqueue_forward(J, st, k == 0, t0, t1)) return -1;
    if(enqueue_t1_blocks(J, st, t0, t1)) return -1;
    if(streamed)
    {
      const uint32_t b0 = J->coded_first[t0], b1 = J->coded_first[t1];
      if(b1 > b0)
      {
        b2k_launch_scan_lengths(J->d_out + b0, J->d_offsets + b0, b1 - b0, st);
        b2k_launch_ht_gather(J->d_enc_desc + b0, J->d_out + b0, J->d_offsets + b0, J->d_scratch, J->d_bytes, b1 - b0,
                             J->bytes_cap, st);
      }
      CUDA_TRY(cudaMemcpyAsync(&J->h_offsets[b1], J->d_offsets + b1, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaEventRecord(J->chunk_ev[CEV(2, k)], st));
      enqueued = k + 1;
      if(return_chunks(false)) return -1;
    }
  }
  CUDA_TRY(cudaEventRecord(J->ev[2], st));
  DBG_T("encode: chunks enqueued");
  b2k_result* R = nullptr;
  const uint32_t nb_all = (uint32_t)J->h_enc_desc.size();
  b2k_result* shell = result_shell(J); /* host work while the device finishes the last chunks */
  if(streamed)
  {
    if(return_chunks(true)) return -1;
    CUDA_TRY(cudaEventRecord(J->chunk_ev[CEV(4, 0)], ds));
    CUDA_TRY(cudaStreamWaitEvent(st, J->chunk_ev[CEV(4, 0)], 0));
    DBG_T("encode: last chunk coded");
    J->bytes_used = total;
    if(overflow)
    { /* estimate too small: grow, compact everything again from the scratch slots, plain copy */
      pool_put(hb);
      hb = nullptr;
      CUDA_TRY(cudaStreamSynchronize(st));
      cudaFree(J->d_bytes);
      J->bytes_cap = total + total / 8 + 4096;
      CUDA_TRY(cudaMalloc(&J->d_bytes, J->bytes_cap));
      b2k_launch_ht_gather(J->d_enc_desc, J->d_out, J->d_offsets, J->d_scratch, J->d_bytes, nb_all, J->bytes_cap, st);
      CUDA_TRY(cudaEventRecord(J->ev[3], st));
      if(int frc = fetch_result(J, st, &R, nullptr, shell)) return frc;
    }
    else
    {
      CUDA_TRY(cudaEventRecord(J->ev[3], st));
      uint8_t* owned = hb;
      hb = nullptr; /* fetch_result hands it to the result, or frees it with the result on failure */
      if(int frc = fetch_result(J, st, &R, owned, shell, true)) return frc;
    }
  }
  if(!streamed)
  {
    b2k_launch_scan_lengths(J->d_out, J->d_offsets, nb_all, st);
    if(finish_t1_encode(J, st)) return -1;
    CUDA_TRY(cudaEventRecord(J->ev[3], st));
    if(int frc = fetch_result(J, st, &R, nullptr, shell)) return frc;
  }
  DBG_T("encode: result fetched");
  CUDA_TRY(cudaEventRecord(J->ev[6], st));
  CUDA_TRY(cudaEventSynchronize(J->ev[6]));
  DBG_T("encode: done");
  float a = 0, b = 0, c = 0, d = 0;
  cudaEventElapsedTime(&a, J->ev[0], J->ev[1]);
  cudaEventElapsedTime(&b, J->ev[1], J->ev[2]);
  cudaEventElapsedTime(&c, J->ev[2], J->ev[3]);
  cudaEventElapsedTime(&d, J->ev[3], J->ev[6]);
  cudaEventElapsedTime(&J->last_level1_ms, J->ev[4], J->ev[5]);
  R->ms_h2d = a; R->ms_dwt = b; R->ms_t1 = c; R->ms_d2h = d; R->ms_total = a + b + c + d;
  if(tuned)
    J->tune_enc.record(pack, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
  *out = R;
  return 0;
}

extern "C" int32_t b2k_encode(b2k_engine* e, const b2k_coding* cp, const int32_t* const* planes, const uint32_t* strides,
                              uint32_t tile_mod, uint32_t tile_rem, b2k_result** out)
{
  return encode_common(e, cp, (void* const*)planes, strides, tile_mod, tile_rem, out, false);
}

extern "C" int32_t b2k_encode16(b2k_engine* e, const b2k_coding* cp, const uint16_t* const* planes, const uint32_t* strides,
                                uint32_t tile_mod, uint32_t tile_rem, b2k_result** out)
{
  return encode_common(e, cp, (void* const*)planes, strides, tile_mod, tile_rem, out, true);
}

extern "C" int32_t b2k_encode16_interleaved(b2k_engine* e, const b2k_coding* cp, const uint16_t* pixels, uint32_t stride,
                                            uint32_t tile_mod, uint32_t tile_rem, b2k_result** out)
{
  if(!cp || !pixels || stride < (uint32_t)(cp->x1 - cp->x0) * cp->numcomps)
    return -1;
  void* planes[4] = {const_cast<uint16_t*>(pixels), nullptr, nullptr, nullptr};
  const uint32_t strides[4] = {stride, 0, 0, 0};
  return encode_common(e, cp, planes, strides, tile_mod, tile_rem, out, true, true);
}

static int32_t decode_common(b2k_engine* e, const b2k_coding* cp, const b2k_block* blocks, uint64_t num_blocks,
                             const uint8_t* bytes, uint64_t num_bytes, void* const* planes, const uint32_t* strides,
                             uint32_t tile_mod, uint32_t tile_rem, double* ms_total, bool u16, const uint32_t* crop = nullptr);

extern "C" int32_t b2k_decode(b2k_engine* e, const b2k_coding* cp, const b2k_block* blocks, uint64_t num_blocks,
                              const uint8_t* bytes, uint64_t num_bytes, int32_t* const* planes, const uint32_t* strides,
                              uint32_t tile_mod, uint32_t tile_rem, double* ms_total)
{
  return decode_common(e, cp, blocks, num_blocks, bytes, num_bytes, (void* const*)planes, strides, tile_mod, tile_rem,
                       ms_total, false);
}
extern "C" int32_t b2k_decode16(b2k_engine* e, const b2k_coding* cp, const b2k_block* blocks, uint64_t num_blocks,
                                const uint8_t* bytes, uint64_t num_bytes, uint16_t* const* planes, const uint32_t* strides,
                                uint32_t tile_mod, uint32_t tile_rem, double* ms_total)
{
  return decode_common(e, cp, blocks, num_blocks, bytes, num_bytes, (void* const*)planes, strides, tile_mod, tile_rem,
                       ms_total, true);
}

/* b2k_decode / b2k_decode16 with only `window` (x0, y0, x1, y1 in cp's canvas coordinates) of the pixels returned: planes[c]
   holds window rows of strides[c] samples.  With b2k_codestream_parse_window this is the windowed decode of SURVEY 8f N3: the
   tiles the window touches are decoded, the window's pixels alone cross PCIe. */
extern "C" int32_t b2k_decode_window(b2k_engine* e, const b2k_coding* cp, const b2k_block* blocks, uint64_t num_blocks,
                                     const uint8_t* bytes, uint64_t num_bytes, void* const* planes, const uint32_t* strides,
                                     const uint32_t* window, uint32_t sample_bytes, double* ms_total)
{
  if(!window || (sample_bytes != 2 && sample_bytes != 4))
    return -1;
  return decode_common(e, cp, blocks, num_blocks, bytes, num_bytes, planes, strides, 1, 0, ms_total, sample_bytes == 2, window);
}

static int32_t decode_common(b2k_engine* e, const b2k_coding* cp, const b2k_block* blocks, uint64_t num_blocks,
                             const uint8_t* bytes, uint64_t num_bytes, void* const* planes, const uint32_t* strides,
                             uint32_t tile_mod, uint32_t tile_rem, double* ms_total, bool u16, const uint32_t* crop)
{
  if(!e || !cp || !blocks || !planes || !strides)
    return -1;
  if(crop && (crop[0] >= crop[2] || crop[1] >= crop[3] || crop[0] < cp->x0 || crop[1] < cp->y0 || crop[2] > cp->x1 || crop[3] > cp->y1))
  {
    g_err = "window outside the image";
    return -1;
  }
  std::lock_guard<std::mutex> lock(e->mu);
  int rc = 0;
  b2k_device_job* J = cached_job(e, cp, tile_mod, tile_rem, &rc);
  if(rc)
    return rc;
  CUDA_TRY(cudaSetDevice(e->device));
  /* a window: only its pixels cross PCIe on the way back, into planes of the window's size (copy_planes / copy_planes16) */
  J->has_crop = crop != nullptr;
  if(crop)
    J->crop = Rect{crop[0], crop[1], crop[2], crop[3]};
  struct CropEnd
  {
    b2k_device_job* j;
    ~CropEnd() { j->has_crop = false; }
  } crop_end{J};
  void* const* user_planes = planes;
  const uint32_t* user_strides = strides;
  void* stage_planes[4];
  uint32_t stage_strides[4];
  /* several ranks on one host share its DRAM and CPU quota: measured packing loses there,
     so the automatic policy only considers it for a process that has the host to itself */
  const bool tuned = !crop && !u16 && host_pack_eligible(J) && g_pack_policy.load() < 0 && b2k_host_local_peers() == 1;
  const bool pack = !crop && !u16 && host_pack_eligible(J) && (tuned ? J->tune_dec.next_mode() : g_pack_policy.load() > 0);
  const auto wall0 = std::chrono::steady_clock::now();
  if(!u16)
    g_last_pack[1].store(pack ? 1 : 0);
  const bool ring = pack && ring_geom(true).slot_mb > 0;
  if(pack)
  {
    if(ring ? ensure_ring(J, true) : ensure_stage16(J)) return -1;
    if(!ring)
    {
      stage16_views(J, stage_planes, stage_strides);
      planes = stage_planes;
      strides = stage_strides;
    }
    u16 = true;
    b2k_host_session(true);
  }
  struct SessionEnd
  {
    bool on;
    ~SessionEnd() { if(on) b2k_host_session(false); }
  } session_end{pack};
  if(u16)
    if(int urc = ensure_u16(J))
      return urc;
  cudaStream_t st = e->stream;
  if(num_bytes + 64 > J->bytes_cap)
  {
    cudaFree(J->d_bytes);
    J->bytes_cap = num_bytes + 4096;
    CUDA_TRY(cudaMalloc(&J->d_bytes, J->bytes_cap));
  }
  CUDA_TRY(cudaEventRecord(J->ev[0], st));
  CUDA_TRY(cudaMemsetAsync(J->d_err, 0, sizeof(int), st));
  J->dec_has_refinement = false;
  cudaStream_t cs = e->copy_stream;
  const size_t nchunks = J->chunk_tile.size() - 1;
  CUDA_TRY(cudaStreamWaitEvent(cs, J->ev[0], 0));
  CUDA_TRY(cudaStreamWaitEvent(e->h2d_stream, J->ev[0], 0));
  /* Pipeline over tile chunks.  Host: build chunk k's descriptors while the device works on chunk
     k-1.  PCIe in: chunk k's coded bytes (one contiguous range when the arena is in block order,
     which ours always is; otherwise the whole arena goes up once).  Compute: block decode +
     inverse transform.  PCIe out: chunk k-1's pixels -- both directions stay busy. */
  bool all_uploaded = false;
  uint64_t prev_end = 0;
  for(size_t k = 0; k < nchunks; ++k)
  {
    const size_t t0 = J->chunk_tile[k], t1 = J->chunk_tile[k + 1];
    const uint32_t b0 = J->coded_first[t0], b1 = J->coded_first[t1];
    /* descriptors travel on the upload stream with the chunk's bytes, so the side-stream parse of
       chunk k+1 depends on nothing the main stream is still doing for chunk k */
    if(int prc = prepare_decode(J, blocks, num_blocks, e->h2d_stream, b0, b1)) return prc;
    uint64_t lo = UINT64_MAX, hi = 0;
    for(uint32_t b = b0; b < b1; ++b)
    {
      const HtBlockDesc& d = J->h_dec_desc[b];
      if(!d.length)
        continue;
      lo = std::min<uint64_t>(lo, d.slot_off);
      hi = std::max<uint64_t>(hi, d.slot_off + d.length + d.length2);
    }
    if(hi > num_bytes)
    {
      g_err = "block offsets exceed the byte arena";
      return -1;
    }
    if(!all_uploaded && lo != UINT64_MAX)
    {
      if(lo < prev_end)
      { /* arena not in block order (a foreign codestream whose tile parts are out of tile order, a caller arena laid
           out some other way): ranges below prev_end that no earlier chunk covered may be needed now, so the whole
           arena goes up once; bytes already on the device are simply written again with the same values */
        CUDA_TRY(cudaMemcpyAsync(J->d_bytes, bytes, num_bytes, cudaMemcpyHostToDevice, e->h2d_stream));
        all_uploaded = true;
      }
      else
      {
        CUDA_TRY(cudaMemcpyAsync(J->d_bytes + lo, bytes + lo, hi - lo, cudaMemcpyHostToDevice, e->h2d_stream));
        prev_end = hi;
      }
    }
    CUDA_TRY(cudaEventRecord(J->chunk_ev[CEV(1, k)], e->h2d_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, J->chunk_ev[CEV(1, k)], 0));
    if(b1 > b0)
    {
      /* phase A (serial VLC/MEL parse, one thread per block) is latency-bound and leaves the SMs
         nearly empty: run the chunks' parses concurrently on side streams, ahead of the main stream */
      cudaStream_t ax = e->aux[k & 3];
      CUDA_TRY(cudaStreamWaitEvent(ax, J->chunk_ev[CEV(1, k)], 0)); /* this chunk's descriptors + bytes are up */
      b2k_launch_ht_decode_vlc(J->d_dec_desc + b0, J->d_bytes, J->d_recs, J->d_dec_status + b0, b1 - b0, J->max_cblk_w, ax);
      CUDA_TRY(cudaEventRecord(J->chunk_ev[CEV(2, k)], ax));
      CUDA_TRY(cudaStreamWaitEvent(st, J->chunk_ev[CEV(2, k)], 0));
      b2k_launch_ht_decode_magsgn(J->d_dec_desc + b0, J->d_bytes, J->d_recs, J->d_dec_status + b0, b1 - b0, J->max_cblk_w,
                                  J->d_err, J->cp.irreversible, J->dec_has_refinement, st);
      if(J->dec_has_refinement)
        b2k_launch_ht_decode_refine(J->d_dec_desc + b0, J->d_bytes, J->d_dec_status + b0, b1 - b0,
                                    (J->cp.cblk_sty & 0x08) != 0, st);
    }
    if(enqueue_inverse(J, st, t0, t1)) return -1;
    if(u16 && convert_planes16(J, false, st, t0, t1)) return -1;
    CUDA_TRY(cudaEventRecord(J->chunk_ev[CEV(0, k)], st));
    if(ring)
      continue; /* pixels come down piece by piece below */
    CUDA_TRY(cudaStreamWaitEvent(cs, J->chunk_ev[CEV(0, k)], 0));
    if(u16 ? copy_planes16(J, planes, strides, false, cs, t0, t1) : copy_planes(J, J->img, planes, strides, false, cs, t0, t1))
      return -1;
    if(pack)
      CUDA_TRY(cudaEventRecord(J->chunk_ev[CEV(3, k)], cs));
  }
  DBG_T("decode: chunks enqueued");
  if(ring)
  {
    if(ring_download_all(J, user_planes, user_strides, cs)) return -1;
  }
  else if(pack) /* widen chunk k into the caller's planes while chunk k+1 is still coming down */
    for(size_t k = 0; k < nchunks; ++k)
    {
      CUDA_TRY(cudaEventSynchronize(J->chunk_ev[CEV(3, k)]));
      host_convert_chunk(J, user_planes, user_strides, true, J->chunk_tile[k], J->chunk_tile[k + 1]);
    }
  CUDA_TRY(cudaEventRecord(J->chunk_ev[CEV(0, nchunks)], cs));
  CUDA_TRY(cudaStreamWaitEvent(st, J->chunk_ev[CEV(0, nchunks)], 0));
  CUDA_TRY(cudaEventRecord(J->ev[1], st));
  CUDA_TRY(cudaEventSynchronize(J->ev[1]));
  CUDA_TRY(cudaGetLastError());
  DBG_T("decode: done");
  float t = 0;
  cudaEventElapsedTime(&t, J->ev[0], J->ev[1]);
  if(ms_total) *ms_total = t;
  if(tuned)
    J->tune_dec.record(pack, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
  int herr = 0;
  CUDA_TRY(cudaMemcpy(&herr, J->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if(herr)
  {
    g_err = "HT decoder rejected " + std::to_string(herr) + " block(s)";
    return -2;
  }
  return 0;
}
