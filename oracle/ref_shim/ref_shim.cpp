/*
 * oracle/ref_shim/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" doorway onto the REFERENCE's own hot-path kernels, compiled by oracle/Makefile
 * from the sources where they lie under /root/reference (nothing is copied into this repo):
 *   t1/part15/coding/ojph_block_encoder{,_avx2,_avx512}.cpp, ojph_block_decoder{32,_ssse3,_avx2}.cpp,
 *   ojph_block_common.cpp, t1/part15/others/*.cpp           (the vendored OpenJPH HT block coder)
 *   wavelet/WaveletFwd.cpp + highway/hwy/{targets,per_target,abort}.cc  (grk::dwt53 / grk::dwt97)
 * The output, oracle/_ref/libgrok_ref.so, is (a) what pins oracle/j2k_oracle.c, and (b) the
 * "reference" CPU arm of bench.py.  The multi-level driver below restates the single-thread
 * branch of encode<T,DWT> (WaveletFwd.cpp L1337-1514) around the reference's encode_v/encode_h.
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <thread>
#include <atomic>
#include <mutex>
#include <algorithm>

#include "ojph_arch.h"
#include "ojph_mem.h"
#include "ojph_block_encoder.h"
#include "ojph_block_decoder.h"
#include "TFSingleton.h"
#include "WaveletCommon.h"
#include <hwy/per_target.h>

/* statics the reference defines in scheduling/CodecScheduler.cpp L61-68 (not compiled here) */
std::shared_ptr<tf::Executor> TFSingleton::instance_ = nullptr;
std::mutex TFSingleton::mutex_;
size_t TFSingleton::numThreads_;
thread_local tf::Executor* TFSingleton::tlsExec_ = nullptr;
thread_local std::atomic<tf::Executor*>* TFSingleton::tlsOwnerExec_ = nullptr;
thread_local size_t TFSingleton::tlsNumThreads_ = 0;
thread_local uint32_t TFSingleton::tlsWorkerId_ = 0;
thread_local bool TFSingleton::tlsActive_ = false;

using namespace grk::t1::ojph;
using namespace grk::t1::ojph::local;

typedef void (*enc_fn)(ui32*, ui32, ui32, ui32, ui32, ui32, ui32*, mem_elastic_allocator*, coded_lists*&);
typedef bool (*dec_fn)(ui8*, ui32*, ui32, ui32, ui32, ui32, ui32, ui32, ui32, bool);

static std::once_flag g_once;
static enc_fn g_enc[3];
static dec_fn g_dec[3];
static int g_best_enc = 0, g_best_dec = 0;

static void init_once()
{
  std::call_once(g_once, [] {
    initialize_block_encoder_tables();
    g_enc[0] = ojph_encode_codeblock32;
    g_dec[0] = ojph_decode_codeblock32;
    g_enc[1] = g_enc[2] = nullptr;
    g_dec[1] = g_dec[2] = nullptr;
    int lvl = get_cpu_ext_level();
    /* same ladder as CoderOJPH.cpp L45-93 */
    if(lvl >= X86_CPU_EXT_LEVEL_SSSE3) { g_dec[1] = ojph_decode_codeblock_ssse3; g_best_dec = 1; }
    if(lvl >= X86_CPU_EXT_LEVEL_AVX2)
    {
      initialize_block_encoder_tables_avx2();
      g_enc[1] = ojph_encode_codeblock_avx2; g_best_enc = 1;
      g_dec[2] = ojph_decode_codeblock_avx2; g_best_dec = 2;
    }
    if(lvl >= X86_CPU_EXT_LEVEL_AVX512)
    {
      initialize_block_encoder_tables_avx512();
      g_enc[2] = ojph_encode_codeblock_avx512; g_best_enc = 2;
    }
  });
}

extern "C" {

__attribute__((visibility("default"))) int ref_cpu_level(void) { init_once(); return get_cpu_ext_level(); }
__attribute__((visibility("default"))) int ref_best_encoder(void) { init_once(); return g_best_enc; }
__attribute__((visibility("default"))) int ref_best_decoder(void) { init_once(); return g_best_dec; }

/* variant: 0 generic, 1 avx2, 2 avx512, -1 = what the reference would dispatch to here.
 * returns coded length, -2 if the variant is unavailable on this CPU. */
__attribute__((visibility("default"))) int ref_ht_encode(int variant, const uint32_t* buf, uint32_t missing_msbs,
                                                         uint32_t w, uint32_t h, uint32_t stride,
                                                         uint8_t* out, uint32_t cap)
{
  init_once();
  if(variant < 0) variant = g_best_enc;
  if(variant > 2 || !g_enc[variant]) return -2;
  thread_local mem_elastic_allocator* elastic = new mem_elastic_allocator(1048576);
  coded_lists* coded = nullptr;
  ui32 lengths[2] = {0, 0};
  elastic->restart();
  g_enc[variant]((ui32*)buf, missing_msbs, 1, w, h, stride, lengths, elastic, coded);
  if(lengths[0] > cap) return -1;
  memcpy(out, coded->buf, lengths[0]);
  return (int)lengths[0];
}

/* variant: 0 generic32, 1 ssse3, 2 avx2, -1 dispatch.  data must be readable 16 bytes before
 * and after (the caller pads, as CoderOJPH.cpp L212-262 does).  returns 0 ok, -1 fail. */
__attribute__((visibility("default"))) int ref_ht_decode(int variant, uint8_t* data, uint32_t* out, uint32_t missing_msbs,
                                                         uint32_t num_passes, uint32_t len1, uint32_t len2,
                                                         uint32_t w, uint32_t h, uint32_t stride)
{
  init_once();
  if(variant < 0) variant = g_best_dec;
  if(variant > 2 || !g_dec[variant]) return -2;
  return g_dec[variant](data, out, missing_msbs, num_passes, len1, len2, w, h, stride, false) ? 0 : -1;
}

/* same, with the stripe-causal flag of the code-block style (CoderOJPH.cpp L248) */
__attribute__((visibility("default"))) int ref_ht_decode_vsc(int variant, uint8_t* data, uint32_t* out, uint32_t missing_msbs,
                                                             uint32_t num_passes, uint32_t len1, uint32_t len2,
                                                             uint32_t w, uint32_t h, uint32_t stride, int stripe_causal)
{
  init_once();
  if(variant < 0) variant = g_best_dec;
  if(variant > 2 || !g_dec[variant]) return -2;
  return g_dec[variant](data, out, missing_msbs, num_passes, len1, len2, w, h, stride, stripe_causal != 0) ? 0 : -1;
}

static inline uint32_t cdp2(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a + ((1ull << b) - 1)) >> b); }

__attribute__((visibility("default"))) int ref_dwt_lanes(void) { return (int)(hwy::VectorBytes() / 4); }

} /* extern "C" */

/* forward multi-level 2-D transform of one tile component with the reference kernels.
 * buf rows must be padded to a multiple of the SIMD width (the vertical kernels load full
 * vectors); scratch is allocated here. */
template<typename T, typename DWT>
static void ref_fwd_2d(T* buf, uint32_t stride, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, int numres,
                       T dcshift, bool intInput)
{
  const uint32_t lanes = (uint32_t)(hwy::VectorBytes() / sizeof(T));
  uint32_t maxdim = (x1 - x0) > (y1 - y0) ? (x1 - x0) : (y1 - y0);
  T* scratch = (T*)aligned_alloc(64, ((size_t)(maxdim + 2) * lanes * sizeof(T) + 63) & ~(size_t)63);
  DWT dwt;
  for(int resno = numres - 1; resno >= 1; --resno)
  {
    bool first = (resno == numres - 1);
    uint32_t n = (uint32_t)(numres - 1 - resno);
    uint32_t rx0 = cdp2(x0, n), ry0 = cdp2(y0, n), rx1 = cdp2(x1, n), ry1 = cdp2(y1, n);
    uint32_t rw = rx1 - rx0, rh = ry1 - ry0;
    uint8_t px = rx0 & 1, py = ry0 & 1;
    T shift = first ? dcshift : T(0);
    bool ii = std::is_floating_point<T>::value ? (first && intInput) : (first ? intInput : true);
    uint32_t j;
    for(j = 0; j + lanes - 1 < rw; j += lanes)
      dwt.encode_v(buf + j, scratch, rh, py, stride, lanes, shift, ii);
    if(j < rw)
      dwt.encode_v(buf + j, scratch, rh, py, stride, rw - j, shift, ii);
    for(j = 0; j + lanes - 1 < rh; j += lanes)
      dwt.encode_h(buf + (size_t)j * stride, scratch, rw, px, stride, lanes);
    if(j < rh)
      dwt.encode_h(buf + (size_t)j * stride, scratch, rw, px, stride, rh - j);
  }
  free(scratch);
}

extern "C" {
__attribute__((visibility("default"))) void ref_dwt53_fwd_2d(int32_t* buf, uint32_t stride, uint32_t x0, uint32_t y0,
                                                             uint32_t x1, uint32_t y1, int numres, int32_t dcshift)
{
  ref_fwd_2d<int32_t, grk::dwt53>(buf, stride, x0, y0, x1, y1, numres, dcshift, false);
}
__attribute__((visibility("default"))) void ref_dwt97_fwd_2d(float* buf, uint32_t stride, uint32_t x0, uint32_t y0,
                                                             uint32_t x1, uint32_t y1, int numres, float dcshift,
                                                             int int_input)
{
  ref_fwd_2d<float, grk::dwt97>(buf, stride, x0, y0, x1, y1, numres, dcshift, int_input != 0);
}

} /* extern "C" */

/* =============================================================================================
 * CPU baseline driver (bench.py --impl reference / cpu_baseline): the reference's kernels over
 * whole tiles on all host threads, tiles handed out by an atomic counter the way the reference
 * hands them to Taskflow workers (CodeStreamCompress.cpp L943-966; block list built like
 * CompressScheduler.cpp L84-139).  What runs per tile:
 *   encode: row copy into an aligned tile buffer (TileProcessorCompress.cpp L176-240),
 *           DC shift + RCT (restated from mct.cpp L497-531 -- that TU needs the Tile graph),
 *           grk::dwt53 multi-level forward (the reference's code), per block the T1 pre-pass
 *           (restated from CoderOJPH.cpp L121-160) + the reference's dispatched HT encoder;
 *   decode: the reference's dispatched HT decoder + shift filter (PostDecodeFiltersOJPH.h L48-66),
 *           inverse RCT (restated from mct.cpp L201-256); the inverse DWT is timed through the
 *           reference's own single-thread hook grk_bench_dwt_53 (WaveletReverse.h L111-118) on
 *           every worker, once per tile component.
 * ========================================================================================== */
struct ref_block_desc
{
  uint32_t comp, buf_x, buf_y, w, h, kmax;
};

extern "C" double grk_bench_dwt_53(uint32_t width, uint32_t height, uint8_t numres, uint32_t iters);

#include <chrono>
static double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

/* planes: ncomp image planes (int32, row stride `stride`), the first ntiles tiles of a tw x th
 * grid with `tiles_per_row` tiles per row are processed.  coded: per (tile, block) output slots of
 * `slot` bytes; lengths[tile*nblocks + b].  Returns wall seconds. */
__attribute__((visibility("default"))) double ref_bench_encode(const int32_t* const* planes, uint32_t stride, int ncomp,
                                                              int ntiles, int tiles_per_row, uint32_t tw, uint32_t th,
                                                              int prec, int numres, const ref_block_desc* blocks,
                                                              int nblocks, uint8_t* coded, uint32_t slot,
                                                              uint32_t* lengths, int nthreads)
{
  init_once();
  std::atomic<int> next{0};
  const uint32_t tstride = ((tw + 15) / 16) * 16 + 16;
  const int32_t dc = -(1 << (prec - 1));
  auto worker = [&]() {
    std::vector<int32_t*> buf(ncomp);
    for(int c = 0; c < ncomp; ++c)
      buf[c] = (int32_t*)aligned_alloc(64, (size_t)tstride * (th + 2) * sizeof(int32_t));
    std::vector<uint32_t> sm(4096 + 64);
    for(;;)
    {
      const int t = next.fetch_add(1);
      if(t >= ntiles)
        break;
      const uint32_t ox = (uint32_t)(t % tiles_per_row) * tw, oy = (uint32_t)(t / tiles_per_row) * th;
      for(int c = 0; c < ncomp; ++c)
        for(uint32_t y = 0; y < th; ++y)
          memcpy(buf[c] + (size_t)y * tstride, planes[c] + (size_t)(oy + y) * stride + ox, tw * sizeof(int32_t));
      if(ncomp >= 3)
      {
        for(uint32_t y = 0; y < th; ++y)
        {
          int32_t *r = buf[0] + (size_t)y * tstride, *g = buf[1] + (size_t)y * tstride, *b = buf[2] + (size_t)y * tstride;
          for(uint32_t x = 0; x < tw; ++x)
          {
            const int32_t rr = r[x] + dc, gg = g[x] + dc, bb = b[x] + dc;
            r[x] = ((gg + gg) + bb + rr) >> 2;
            g[x] = bb - gg;
            b[x] = rr - gg;
          }
        }
      }
      for(int c = 0; c < ncomp; ++c)
        ref_fwd_2d<int32_t, grk::dwt53>(buf[c], tstride, ox, oy, ox + tw, oy + th, numres, (ncomp >= 3 && c < 3) ? 0 : -dc, false);
      for(int k = 0; k < nblocks; ++k)
      {
        const ref_block_desc& B = blocks[k];
        const int shift = 31 - (int)(B.kmax + 1);
        for(uint32_t y = 0; y < B.h; ++y)
        {
          const int32_t* src = buf[B.comp] + (size_t)(B.buf_y + y) * tstride + B.buf_x;
          for(uint32_t x = 0; x < B.w; ++x)
          {
            const int32_t v = src[x];
            const uint32_t mag = v >= 0 ? (uint32_t)v : (uint32_t)0 - (uint32_t)v;
            sm[y * B.w + x] = (v >= 0 ? 0u : 0x80000000u) | (mag << shift);
          }
        }
        const int n = ref_ht_encode(-1, sm.data(), B.kmax, B.w, B.h, B.w, coded + ((size_t)t * nblocks + k) * slot, slot);
        lengths[(size_t)t * nblocks + k] = n > 0 ? (uint32_t)n : 0;
      }
    }
    for(int c = 0; c < ncomp; ++c)
      free(buf[c]);
  };
  const double t0 = now_s();
  std::vector<std::thread> th_;
  for(int i = 0; i < nthreads; ++i)
    th_.emplace_back(worker);
  for(auto& t : th_)
    t.join();
  return now_s() - t0;
}

__attribute__((visibility("default"))) double ref_bench_decode(int ncomp, int ntiles, uint32_t tw, uint32_t th, int prec,
                                                              int numres, const ref_block_desc* blocks, int nblocks,
                                                              const uint8_t* coded, uint32_t slot, const uint32_t* lengths,
                                                              int nthreads, double* dwt_seconds_per_tilecomp)
{
  init_once();
  std::atomic<int> next{0};
  const uint32_t tstride = ((tw + 15) / 16) * 16 + 16;
  const int32_t dc = 1 << (prec - 1), hi = (1 << prec) - 1;
  std::vector<double> dwt_best(nthreads, 0.0), excess(nthreads, 0.0);
  auto worker = [&](int wid) {
    std::vector<int32_t*> buf(ncomp);
    for(int c = 0; c < ncomp; ++c)
      buf[c] = (int32_t*)aligned_alloc(64, (size_t)tstride * (th + 2) * sizeof(int32_t));
    std::vector<uint32_t> dec((size_t)72 * 72 + 64);
    std::vector<uint8_t> pad(slot + 64);
    int done = 0;
    for(;;)
    {
      const int t = next.fetch_add(1);
      if(t >= ntiles)
        break;
      for(int k = 0; k < nblocks; ++k)
      {
        const ref_block_desc& B = blocks[k];
        const uint32_t len = lengths[(size_t)t * nblocks + k];
        const uint32_t dstride = (B.w + 7u) & ~7u;
        memset(pad.data(), 0, 16);
        memcpy(pad.data() + 16, coded + ((size_t)t * nblocks + k) * slot, len);
        memset(pad.data() + 16 + len, 0, 16);
        ref_ht_decode(-1, pad.data() + 16, dec.data(), B.kmax - 1, 1, len, 0, B.w, B.h, dstride);
        const uint32_t sh = 31u - B.kmax;
        for(uint32_t y = 0; y < B.h; ++y)
        {
          int32_t* dst = buf[B.comp] + (size_t)(B.buf_y + y) * tstride + B.buf_x;
          for(uint32_t x = 0; x < B.w; ++x)
          {
            const uint32_t v = dec[y * dstride + x];
            const int32_t m = (int32_t)((v & 0x7FFFFFFFu) >> sh);
            dst[x] = (v & 0x80000000u) ? -m : m;
          }
        }
      }
      /* inverse DWT: the reference's own hook, once per tile component */
      const double h0 = now_s();
      const double per = grk_bench_dwt_53(tw, th, (uint8_t)numres, (uint32_t)ncomp);
      /* the hook also allocates and fills its buffers: only ncomp * per belongs to a decode */
      excess[wid] += (now_s() - h0) - (double)ncomp * (per > 0 ? per : 0);
      if(per > 0 && (dwt_best[wid] == 0.0 || per < dwt_best[wid]))
        dwt_best[wid] = per;
      if(ncomp >= 3)
        for(uint32_t y = 0; y < th; ++y)
        {
          int32_t *a = buf[0] + (size_t)y * tstride, *b = buf[1] + (size_t)y * tstride, *c = buf[2] + (size_t)y * tstride;
          for(uint32_t x = 0; x < tw; ++x)
          {
            const int32_t g = a[x] - ((b[x] + c[x]) >> 2), r = c[x] + g, bl = b[x] + g;
            a[x] = std::min(std::max(r + dc, 0), hi);
            b[x] = std::min(std::max(g + dc, 0), hi);
            c[x] = std::min(std::max(bl + dc, 0), hi);
          }
        }
      ++done;
    }
    for(int c = 0; c < ncomp; ++c)
      free(buf[c]);
  };
  const double t0 = now_s();
  std::vector<std::thread> th_;
  for(int i = 0; i < nthreads; ++i)
    th_.emplace_back(worker, i);
  for(auto& t : th_)
    t.join();
  double el = now_s() - t0;
  {
    double ex = 0;
    for(double v : excess) ex += v;
    el -= ex / nthreads;
  }
  if(dwt_seconds_per_tilecomp)
  {
    double s = 0;
    int n = 0;
    for(double v : dwt_best)
      if(v > 0) { s += v; ++n; }
    *dwt_seconds_per_tilecomp = n ? s / n : 0.0;
  }
  return el;
}

} /* extern "C" */
