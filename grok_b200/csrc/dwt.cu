/*
 * grok_b200/csrc/dwt.cu -- forward / inverse lifting DWT (reversible 5/3, irreversible 9/7) for
 * sm_100a, one decomposition level per launch, with the DC level shift and the RCT / ICT
 * multi-component transform fused into the finest level.
 *
 * What it replaces in the reference (CPU, Highway SIMD + Taskflow):
 *   forward : Mct::compress_rev / compress_irrev      point_transform/mct.cpp L497-531, L584-636
 *             encode_53_v/h, encode_97_v/h, encode<>  wavelet/WaveletFwd.cpp L139-876, L1337-1514
 *   inverse : tile_53 / tile_97                       wavelet/WaveletReverse.cpp L1347-1397,
 *                                                     wavelet/WaveletReverse97.cpp L837-857, L950-
 *             DecompressRev / DecompressIrrev         point_transform/mct.cpp L201-256, L318-391
 *
 * B200 design (not a port of the column-strip SIMD loops):
 *   - one warp = one job = (tile component(s), 8*strip_w-column strip, row segment).  Each lane
 *     owns 8 consecutive canvas columns (two 128-bit loads per row).  The reference's
 *     "all columns, then all rows" double pass is fused: rows stream through a register
 *     sliding window for the vertical lifting, every finished vertical row is lifted
 *     horizontally with warp shuffles (predict/update need one neighbour each), and the four
 *     sub-bands are written once, de-interleaved, with 128-bit stores.  HBM traffic per level is
 *     one read + one write of every coefficient: the algorithmic minimum.
 *   - symmetric extension is implemented in the LOADS (mirrored addresses), so the lifting
 *     code has no boundary cases; lanes 0 and 31 are halo lanes (recompute instead of
 *     cross-warp exchange), row segments recompute 3 (5/3) or 7 (9/7) halo rows.
 *   - integer maths is bit-exact with the reference; 9/7 uses fmaf() where the reference build
 *     contracts to FMA (see oracle/j2k_oracle.c fwd97_line) and explicit _rn intrinsics elsewhere.
 */
#include "b2k_internal.h"

namespace {

__device__ __noinline__ int mirror_rel_slow(int t, int n)
{
  if(n == 1)
    return 0;
  const int period = 2 * (n - 1);
  t %= period;
  if(t < 0)
    t += period;
  return t < n ? t : period - t;
}
/* whole-sample symmetric extension index; one reflection covers every halo unless the line is
   shorter than the halo, which takes the (out-of-line) periodic path */
__device__ __forceinline__ int mirror_rel(int t, int n)
{
  if((unsigned)t < (unsigned)n)
    return t;
  const int r = t < 0 ? -t : 2 * (n - 1) - t;
  if((unsigned)r < (unsigned)n)
    return r;
  return mirror_rel_slow(t, n);
}

/* ---- 8-sample row loads ------------------------------------------------------------------ */
/* row points at the sample with relative index 0 (canvas u0); rel = first wanted index */
__device__ __forceinline__ void load8_w32(const int32_t* __restrict__ row, int rel, int n, int (&v)[8])
{
  const int32_t* p = row + rel;
  if(rel >= 0 && rel + 8 <= n && ((reinterpret_cast<uintptr_t>(p) & 15) == 0))
  {
    const int4 a = __ldg(reinterpret_cast<const int4*>(p));
    const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  else
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
      v[i] = __ldg(row + mirror_rel(rel + i, n));
  }
}
__device__ __forceinline__ void load8_u16(const uint16_t* __restrict__ row, int rel, int n, int sgnd, int (&v)[8])
{
  const uint16_t* p = row + rel;
  if(rel >= 0 && rel + 8 <= n && ((reinterpret_cast<uintptr_t>(p) & 15) == 0))
  {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = a.x & 0xFFFF; v[1] = a.x >> 16; v[2] = a.y & 0xFFFF; v[3] = a.y >> 16;
    v[4] = a.z & 0xFFFF; v[5] = a.z >> 16; v[6] = a.w & 0xFFFF; v[7] = a.w >> 16;
  }
  else
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
      v[i] = __ldg(row + mirror_rel(rel + i, n));
  }
  if(sgnd)
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
      v[i] = (int)(int16_t)v[i];
  }
}

/* 4 consecutive band samples (inverse transform): idx0 = first band-relative index wanted,
 * mirrored per element through the interleaved domain when out of range */
__device__ __forceinline__ void load4_band(const int32_t* __restrict__ row, int k0, int kb0, int u0, int u1, int odd,
                                           int (&v)[4])
{
  /* sample i has canvas position u = 2*(k0+i)+odd ; band index = (u' >> 1) - kb0 */
  const int ufirst = 2 * k0 + odd, ulast = ufirst + 6;
  const int32_t* p = row + (k0 - kb0);
  if(ufirst >= u0 && ulast < u1 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0))
  {
    const int4 a = __ldg(reinterpret_cast<const int4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  else
  {
    const int n = u1 - u0;
#pragma unroll
    for(int i = 0; i < 4; ++i)
    {
      const int um = u0 + mirror_rel(ufirst + 2 * i - u0, n);
      /* a length-1 line mirrors onto a sample of the other parity: that band is empty */
      v[i] = ((um & 1) == odd) ? __ldg(row + ((um >> 1) - kb0)) : 0;
    }
  }
}

__device__ __forceinline__ void store4(int32_t* __restrict__ row, int col, const int (&v)[4], unsigned validmask)
{
  int32_t* p = row + col;
  if(validmask == 0xF && ((reinterpret_cast<uintptr_t>(p) & 15) == 0))
    *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]);
  else
  {
#pragma unroll
    for(int i = 0; i < 4; ++i)
      if(validmask & (1u << i))
        p[i] = v[i];
  }
}
__device__ __forceinline__ void store8(int32_t* __restrict__ row, int col, const int (&v)[8], unsigned validmask)
{
  int32_t* p = row + col;
  if(validmask == 0xFF && ((reinterpret_cast<uintptr_t>(p) & 15) == 0))
  {
    reinterpret_cast<int4*>(p)[0] = make_int4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<int4*>(p)[1] = make_int4(v[4], v[5], v[6], v[7]);
  }
  else
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
      if(validmask & (1u << i))
        p[i] = v[i];
  }
}

/* ---- job decoding ------------------------------------------------------------------------- */
struct Job
{
  int lane, ulane, jbeg, jend, wn, hn, nvalid;
  bool owner, need;
};
__device__ __forceinline__ bool decode_job(const DwtLevelDesc& D, Job& J)
{
  J.lane = threadIdx.x & 31;
  const int job = blockIdx.x * B2K_WARPS_PER_CTA + (threadIdx.x >> 5);
  if(job >= (int)D.nstrips * (int)D.nsegs)
    return false;
  const int strip = job % D.nstrips, seg = job / D.nstrips;
  J.wn = D.u1 - D.u0;
  J.hn = D.v1 - D.v0;
  J.nvalid = D.strip_w >> 3;
  J.ulane = (D.u0 & ~7) + strip * (int)D.strip_w + (J.lane - 1) * 8;
  J.owner = J.lane >= 1 && J.lane <= J.nvalid && J.ulane < D.u1;
  J.need = J.lane <= J.nvalid + 1 && J.ulane < D.u1 + 8;
  const int jlo = D.v0 >> 1, jhi = (D.v1 - 1) >> 1;
  J.jbeg = jlo + seg * (int)D.pairs_per_seg;
  J.jend = min(J.jbeg + (int)D.pairs_per_seg, jhi + 1);
  return true;
}

/* =============================================================================================
 * forward, finest-level sample fetch: integers from the image, + DC shift, + RCT / ICT
 * =========================================================================================== */
template <int NC, bool U16>
__device__ __forceinline__ void fetch_int_rows(const DwtLevelDesc& D, const Job& J, int v, int (&out)[NC][8])
{
  const int r = mirror_rel(v - D.v0, J.hn);
#pragma unroll
  for(int c = 0; c < NC; ++c)
  {
    if(!J.need)
    {
#pragma unroll
      for(int i = 0; i < 8; ++i)
        out[c][i] = 0;
      continue;
    }
    if(U16)
      load8_u16(reinterpret_cast<const uint16_t*>(D.in[c]) + (size_t)r * D.in_pitch, J.ulane - D.u0, J.wn,
                D.in_is_u16 == 2, out[c]);
    else
      load8_w32(reinterpret_cast<const int32_t*>(D.in[c]) + (size_t)r * D.in_pitch, J.ulane - D.u0, J.wn, out[c]);
  }
}

/* reversible: mct.cpp L497-531 */
template <int NC, bool U16>
__device__ __forceinline__ void fetch53(const DwtLevelDesc& D, const Job& J, int v, int (&out)[NC][8])
{
  if(D.first_level)
  {
    fetch_int_rows<NC, U16>(D, J, v, out);
#pragma unroll
    for(int i = 0; i < 8; ++i)
    {
      if(NC == 3)
      {
        const int r = out[0][i] + D.shift[0], g = out[1][i] + D.shift[1], b = out[2][i] + D.shift[2];
        out[0][i] = ((g + g) + b + r) >> 2;
        out[1][i] = b - g;
        out[2][i] = r - g;
      }
      else
        out[0][i] += D.shift[0];
    }
  }
  else
    fetch_int_rows<NC, false>(D, J, v, out);
}

/* irreversible: mct.cpp L584-636; float conversion of the finest level WaveletFwd.cpp L658-681 */
template <int NC, bool U16>
__device__ __forceinline__ void fetch97(const DwtLevelDesc& D, const Job& J, int v, float (&out)[NC][8])
{
  int raw[NC][8];
  if(D.first_level)
  {
    fetch_int_rows<NC, U16>(D, J, v, raw);
    const float a_r = 0.299f, a_g = 0.587f, a_b = 0.114f;
    const float cb = __fdiv_rn(0.5f, __fsub_rn(1.0f, a_b)), cr = __fdiv_rn(0.5f, __fsub_rn(1.0f, a_r));
#pragma unroll
    for(int i = 0; i < 8; ++i)
    {
      if(NC == 3)
      {
        const float r = (float)(raw[0][i] + D.shift[0]), g = (float)(raw[1][i] + D.shift[1]),
                    b = (float)(raw[2][i] + D.shift[2]);
        /* the reference build contracts a_r*r + a_g*g + a_b*b into two FMAs (pinned against libgrokj2k, tests/test_interop.py) */
        const float y = __fmaf_rn(a_b, b, __fmaf_rn(a_g, g, __fmul_rn(a_r, r)));
        out[0][i] = y;
        out[1][i] = __fmul_rn(cb, __fsub_rn(b, y));
        out[2][i] = __fmul_rn(cr, __fsub_rn(r, y));
      }
      else
        out[0][i] = (float)(raw[0][i] + D.shift[0]);
    }
  }
  else
  {
    fetch_int_rows<NC, false>(D, J, v, raw);
#pragma unroll
    for(int c = 0; c < NC; ++c)
#pragma unroll
      for(int i = 0; i < 8; ++i)
        out[c][i] = __int_as_float(raw[c][i]);
  }
}

/* ---- sub-band row stores (Mallat layout: TileComponentWindow.h L241-264) -------------------- */
struct BandGeom
{
  int x0l, x0h, y0l, y0h, snx, sny;
};
__device__ __forceinline__ BandGeom band_geom(const DwtLevelDesc& D)
{
  BandGeom g;
  g.x0l = (D.u0 + 1) >> 1;
  g.x0h = D.u0 >> 1;
  g.y0l = (D.v0 + 1) >> 1;
  g.y0h = D.v0 >> 1;
  g.snx = ((D.u1 + 1) >> 1) - g.x0l;
  g.sny = ((D.v1 + 1) >> 1) - g.y0l;
  return g;
}

/* lo[4]/hi[4]: horizontally transformed samples of one vertical row (vertical low if !vhigh) */
__device__ __forceinline__ void store_band_rows(const DwtLevelDesc& D, const Job& J, const BandGeom& g, int c, int j,
                                                bool vhigh, const int (&lo)[4], const int (&hi)[4])
{
  const int v = 2 * j + (vhigh ? 1 : 0);
  if(!J.owner || v < D.v0 || v >= D.v1)
    return;
  unsigned mlo = 0, mhi = 0;
#pragma unroll
  for(int i = 0; i < 4; ++i)
  {
    const int ue = J.ulane + 2 * i;
    if(ue >= D.u0 && ue < D.u1)
      mlo |= 1u << i;
    if(ue + 1 >= D.u0 && ue + 1 < D.u1)
      mhi |= 1u << i;
  }
  const int k0 = J.ulane >> 1;
  if(!vhigh)
  {
    /* LL -> out_ll, HL -> Mallat top-right */
    int32_t* llrow = reinterpret_cast<int32_t*>(D.out_ll[c]) + (size_t)(j - g.y0l) * D.ll_pitch;
    store4(llrow, k0 - g.x0l, lo, mlo);
    int32_t* crow = reinterpret_cast<int32_t*>(D.out_c[c]) + (size_t)(j - g.y0l) * D.c_pitch;
    store4(crow, g.snx + k0 - g.x0h, hi, mhi);
  }
  else
  {
    int32_t* crow = reinterpret_cast<int32_t*>(D.out_c[c]) + (size_t)(g.sny + j - g.y0h) * D.c_pitch;
    store4(crow, k0 - g.x0l, lo, mlo);
    store4(crow, g.snx + k0 - g.x0h, hi, mhi);
  }
}

/* ---- horizontal lifting of one row held as 8 values per lane -------------------------------- */
template <bool DEGEN = true>
__device__ __forceinline__ void hfwd53(const int (&r)[8], int wn, int (&lo)[4], int (&hi)[4])
{
  const int en = __shfl_down_sync(0xffffffffu, r[0], 1);
  hi[0] = r[1] - ((r[0] + r[2]) >> 1);
  hi[1] = r[3] - ((r[2] + r[4]) >> 1);
  hi[2] = r[5] - ((r[4] + r[6]) >> 1);
  hi[3] = r[7] - ((r[6] + en) >> 1);
  const int dp = __shfl_up_sync(0xffffffffu, hi[3], 1);
  lo[0] = r[0] + ((dp + hi[0] + 2) >> 2);
  lo[1] = r[2] + ((hi[0] + hi[1] + 2) >> 2);
  lo[2] = r[4] + ((hi[1] + hi[2] + 2) >> 2);
  lo[3] = r[6] + ((hi[2] + hi[3] + 2) >> 2);
  if(DEGEN && wn == 1)
  { /* WaveletFwd.cpp L289-300: lone column, doubled when it sits on an odd coordinate */
#pragma unroll
    for(int i = 0; i < 4; ++i)
    {
      lo[i] = r[2 * i];
      hi[i] = r[2 * i + 1] << 1;
    }
  }
}

#define F97_ALPHA (-1.586134342f)
#define F97_BETA (-0.052980118f)
#define F97_GAMMA (0.882911075f)
#define F97_DELTA (0.443506852f)
#define F97_K (1.230174105f)

__device__ __forceinline__ void hfwd97(const float (&r)[8], int wn, float invK, float deltaS, int (&lo)[4],
                                       int (&hi)[4])
{
  float e[5], d[4], s[5];
#pragma unroll
  for(int i = 0; i < 4; ++i)
    e[i] = r[2 * i];
  e[4] = __shfl_down_sync(0xffffffffu, r[0], 1);
#pragma unroll
  for(int i = 0; i < 4; ++i)
    d[i] = fmaf(e[i] + e[i + 1], F97_ALPHA, r[2 * i + 1]);
  float dm = __shfl_up_sync(0xffffffffu, d[3], 1);
  s[0] = fmaf(dm + d[0], F97_BETA, e[0]);
#pragma unroll
  for(int i = 1; i < 4; ++i)
    s[i] = fmaf(d[i - 1] + d[i], F97_BETA, e[i]);
  s[4] = __shfl_down_sync(0xffffffffu, s[0], 1);
#pragma unroll
  for(int i = 0; i < 4; ++i)
    d[i] = __fmul_rn(fmaf(s[i] + s[i + 1], F97_GAMMA, d[i]), F97_K);
  dm = __shfl_up_sync(0xffffffffu, d[3], 1);
  float o0 = __fmul_rn(fmaf(dm + d[0], deltaS, s[0]), invK);
  lo[0] = __float_as_int(o0);
#pragma unroll
  for(int i = 1; i < 4; ++i)
    lo[i] = __float_as_int(__fmul_rn(fmaf(d[i - 1] + d[i], deltaS, s[i]), invK));
#pragma unroll
  for(int i = 0; i < 4; ++i)
    hi[i] = __float_as_int(d[i]);
  if(wn == 1)
  { /* WaveletFwd.cpp L444-455 */
#pragma unroll
    for(int i = 0; i < 4; ++i)
    {
      lo[i] = __float_as_int(r[2 * i]);
      hi[i] = __float_as_int(__fmul_rn(r[2 * i + 1], 2.0f));
    }
  }
}

/* =============================================================================================
 * staged row fetch for the forward kernels: every lane prefetches ITS OWN 8 samples of the next
 * row pairs into a private shared-memory slot with cp.async (LDGSTS, 16 bytes per copy, L1
 * bypass) and reads them back later, so STAGES-1 row pairs (x NC components) are in flight per
 * warp without holding registers.  No cross-lane traffic -> no barrier; edge lanes (mirrored or
 * unaligned columns) fill their slot with ordinary loads.
 * =========================================================================================== */
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

/* ---- bulk asynchronous copies (the TMA engine: SASS UBLKCP) completing on an mbarrier -------------------------
 * A warp whose 32 lanes all sit inside the line stages a whole 1 KB warp-row with ONE instruction issued by one lane
 * instead of 64 LDGSTS (two per lane): the copy engine generates the addresses, the LSU issue slots go back to the
 * lifting code.  The row lands linearly (no XOR swizzle -- a bulk copy is contiguous); the lanes' two 128-bit reads
 * are then 2-way bank conflicted, which costs nothing here: shared memory moves 16 B/clk/SM of the 128 it can. */
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "MBAR_WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra MBAR_DONE_%=;\n"
               "bra MBAR_WAIT_%=;\n"
               "MBAR_DONE_%=:\n"
               "}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(parity)
               : "memory");
}

template <int NC, bool U16>
struct RowStage
{
  static constexpr int ROWB = U16 ? 512 : 1024; /* bytes of one warp-row of one component */
  static constexpr int PAIRB = 2 * NC * ROWB;   /* one row pair, all components */
  /* 16-byte chunk k of a 32-bit warp-row (lane L owns chunks 2L, 2L+1) is stored at chunk slot
     k ^ ((k >> 3) & 1): the two 128-bit reads of a lane then hit disjoint banks per quarter warp */
  static __device__ __forceinline__ int swz(int k) { return k ^ ((k >> 3) & 1); }

  /* fill slot `which` (0 = odd row, 1 = next even row) of a stage with canvas row v.
     32-bit rows are copied COOPERATIVELY: one cp.async instruction moves 512 contiguous bytes
     (lane L copies chunks L and L+32), so every 32-byte DRAM sector is requested once.
     fastmask: ballot of the lanes whose 8 columns are interior and aligned. */
  static __device__ __forceinline__ void fill(uint8_t* stage, int which, const DwtLevelDesc& D, const Job& J, int v,
                                              bool fast, unsigned fastmask, const int (&mcol)[8])
  {
    const int r = mirror_rel(v - D.v0, J.hn);
    const int rel = J.ulane - D.u0;
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      uint8_t* dst = stage + (which * NC + c) * ROWB;
      if(U16)
      {
        const uint16_t* row = reinterpret_cast<const uint16_t*>(D.in[c]) + (size_t)r * D.in_pitch;
        if(fast)
          cp_async16(dst + J.lane * 16, row + rel);
        else if(J.need)
        {
          uint32_t w[4];
#pragma unroll
          for(int i = 0; i < 4; ++i)
            w[i] = (uint32_t)__ldg(row + mcol[2 * i]) | ((uint32_t)__ldg(row + mcol[2 * i + 1]) << 16);
          *reinterpret_cast<uint4*>(dst + J.lane * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      else
      {
        const int32_t* row = reinterpret_cast<const int32_t*>(D.in[c]) + (size_t)r * D.in_pitch;
        const int rel0 = rel - 8 * J.lane; /* lane 0's first column */
#pragma unroll
        for(int h = 0; h < 2; ++h)
        {
          const int k = J.lane + 32 * h;
          if((fastmask >> (k >> 1)) & 1u)
            cp_async16(dst + swz(k) * 16, row + rel0 + 4 * k);
        }
        if(J.need && !fast)
        { /* edge lane: mirrored columns, still asynchronous (4-byte copies) */
          uint8_t* d0 = dst + swz(2 * J.lane) * 16;
          uint8_t* d1 = dst + swz(2 * J.lane + 1) * 16;
#pragma unroll
          for(int i = 0; i < 4; ++i)
          {
            cp_async4(d0 + 4 * i, row + mcol[i]);
            cp_async4(d1 + 4 * i, row + mcol[4 + i]);
          }
        }
      }
    }
  }
  static __device__ __forceinline__ void read(const uint8_t* stage, int which, const DwtLevelDesc& D, const Job& J,
                                              int (&out)[NC][8])
  {
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      const uint8_t* src = stage + (which * NC + c) * ROWB;
      if(U16)
      {
        const uint4 a = *reinterpret_cast<const uint4*>(src + J.lane * 16);
        out[c][0] = a.x & 0xFFFF; out[c][1] = a.x >> 16; out[c][2] = a.y & 0xFFFF; out[c][3] = a.y >> 16;
        out[c][4] = a.z & 0xFFFF; out[c][5] = a.z >> 16; out[c][6] = a.w & 0xFFFF; out[c][7] = a.w >> 16;
        if(D.in_is_u16 == 2)
        {
#pragma unroll
          for(int i = 0; i < 8; ++i)
            out[c][i] = (int)(int16_t)out[c][i];
        }
      }
      else
      {
        const int4 a = *reinterpret_cast<const int4*>(src + swz(2 * J.lane) * 16),
                   b = *reinterpret_cast<const int4*>(src + swz(2 * J.lane + 1) * 16);
        out[c][0] = a.x; out[c][1] = a.y; out[c][2] = a.z; out[c][3] = a.w;
        out[c][4] = b.x; out[c][5] = b.y; out[c][6] = b.z; out[c][7] = b.w;
      }
      /* lanes beyond the right halo hold stale shared memory: nothing they compute is stored */
    }
  }
  /* bulk path (32-bit samples, every lane fast): one lane asks the copy engine for the NC whole warp-rows of canvas
     row v; they land linearly in the slot and complete on `bar` (the caller has armed it with expect_tx) */
  static __device__ __forceinline__ void fill_bulk(uint8_t* stage, int which, const DwtLevelDesc& D, const Job& J, int v, uint64_t* bar)
  {
    const int r = mirror_rel(v - D.v0, J.hn);
    const int rel0 = J.ulane - D.u0 - 8 * J.lane; /* lane 0's first column */
#pragma unroll
    for(int c = 0; c < NC; ++c)
      bulk_g2s(stage + (which * NC + c) * ROWB, reinterpret_cast<const int32_t*>(D.in[c]) + (size_t)r * D.in_pitch + rel0, ROWB, bar);
  }
  static __device__ __forceinline__ void read_linear(const uint8_t* stage, int which, const Job& J, int (&out)[NC][8])
  {
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      const uint8_t* src = stage + (which * NC + c) * ROWB + J.lane * 32;
      const int4 a = *reinterpret_cast<const int4*>(src), b = *reinterpret_cast<const int4*>(src + 16);
      out[c][0] = a.x; out[c][1] = a.y; out[c][2] = a.z; out[c][3] = a.w;
      out[c][4] = b.x; out[c][5] = b.y; out[c][6] = b.z; out[c][7] = b.w;
    }
  }
  /* lane can use 16-byte async copies: its 8 columns are inside the line and 16-byte aligned */
  static __device__ __forceinline__ bool lane_fast(const DwtLevelDesc& D, const Job& J)
  {
    const int rel = J.ulane - D.u0;
    bool ok = J.need && rel >= 0 && rel + 8 <= J.wn && ((D.in_pitch * (U16 ? 2u : 4u)) & 15u) == 0;
#pragma unroll
    for(int c = 0; c < NC; ++c)
      ok = ok && (((reinterpret_cast<uintptr_t>(D.in[c]) + (size_t)rel * (U16 ? 2 : 4)) & 15) == 0);
    return ok;
  }
};

/* integer samples of the finest level -> DC shift (+ RCT): mct.cpp L497-531 */
template <int NC>
__device__ __forceinline__ void rct_fwd_inplace(const DwtLevelDesc& D, int (&x)[NC][8])
{
  if(!D.first_level)
    return;
#pragma unroll
  for(int i = 0; i < 8; ++i)
  {
    if(NC == 3)
    {
      const int r = x[0][i] + D.shift[0], g = x[NC > 1 ? 1 : 0][i] + D.shift[1], b = x[NC > 2 ? 2 : 0][i] + D.shift[2];
      x[0][i] = ((g + g) + b + r) >> 2;
      x[NC > 1 ? 1 : 0][i] = b - g;
      x[NC > 2 ? 2 : 0][i] = r - g;
    }
    else
      x[0][i] += D.shift[0];
  }
}

/* per-lane constants of the sub-band stores */
struct StoreCtx
{
  unsigned mlo, mhi;
  bool vec_ll, vec_lo, vec_hi; /* all four samples valid and the 16-byte store is aligned */
  int col_ll, col_lo, col_hi; /* column of the lane's first low sample in the LL plane / in the
                                 Mallat buffer, and of its first high sample */
};
__device__ __forceinline__ StoreCtx store_ctx(const DwtLevelDesc& D, const Job& J, const BandGeom& g)
{
  StoreCtx s;
  s.mlo = s.mhi = 0;
  if(J.owner)
  {
#pragma unroll
    for(int i = 0; i < 4; ++i)
    {
      const int ue = J.ulane + 2 * i;
      if(ue >= D.u0 && ue < D.u1)
        s.mlo |= 1u << i;
      if(ue + 1 >= D.u0 && ue + 1 < D.u1)
        s.mhi |= 1u << i;
    }
  }
  const int k0 = J.ulane >> 1;
  s.col_ll = k0 - g.x0l;
  s.col_lo = k0 - g.x0l;
  s.col_hi = g.snx + k0 - g.x0h;
  /* row pitches are multiples of 4 elements (engine allocates 128-byte multiples) */
  const bool pitch_ok = ((D.ll_pitch | D.c_pitch) & 3u) == 0;
  s.vec_ll = pitch_ok && s.mlo == 0xF && (((reinterpret_cast<uintptr_t>(D.out_ll[0]) >> 2) + (unsigned)s.col_ll) & 3u) == 0;
  s.vec_lo = pitch_ok && s.mlo == 0xF && (((reinterpret_cast<uintptr_t>(D.out_c[0]) >> 2) + (unsigned)s.col_lo) & 3u) == 0;
  s.vec_hi = pitch_ok && s.mhi == 0xF && (((reinterpret_cast<uintptr_t>(D.out_c[0]) >> 2) + (unsigned)s.col_hi) & 3u) == 0;
  return s;
}
__device__ __forceinline__ void store4v(int32_t* __restrict__ p, const int (&v)[4], unsigned mask, bool vec)
{
  if(vec)
    *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]);
  else
  {
#pragma unroll
    for(int i = 0; i < 4; ++i)
      if(mask & (1u << i))
        p[i] = v[i];
  }
}
__device__ __forceinline__ void store_rows_fast(const DwtLevelDesc& D, const BandGeom& g, const StoreCtx& S, int c, int j,
                                                bool vhigh, const int (&lo)[4], const int (&hi)[4])
{
  const int v = 2 * j + (vhigh ? 1 : 0);
  if(v < D.v0 || v >= D.v1 || (S.mlo | S.mhi) == 0)
    return;
  /* all components of a descriptor share the alignment of component 0 (planes are equally laid out) */
  if(!vhigh)
  {
    const int r = j - g.y0l;
    store4v(reinterpret_cast<int32_t*>(D.out_ll[c]) + (r * (int)D.ll_pitch + S.col_ll), lo, S.mlo, S.vec_ll);
    store4v(reinterpret_cast<int32_t*>(D.out_c[c]) + (r * (int)D.c_pitch + S.col_hi), hi, S.mhi, S.vec_hi);
  }
  else
  {
    const int r = g.sny + j - g.y0h;
    int32_t* crow = reinterpret_cast<int32_t*>(D.out_c[c]) + r * (int)D.c_pitch;
    store4v(crow + S.col_lo, lo, S.mlo, S.vec_lo);
    store4v(crow + S.col_hi, hi, S.mhi, S.vec_hi);
  }
}


/* a tile component one sample wide or high at this level: straightforward, unpipelined path
   (WaveletFwd.cpp L146-156, L289-300 special cases live here, not in the hot loop) */
template <int NC, bool U16>
__device__ __noinline__ void fwd53_degenerate_job(const DwtLevelDesc* __restrict__ dptr, const Job J)
{
  const DwtLevelDesc& D = *dptr; /* re-read from global memory: this path is cold */
  const BandGeom g = band_geom(D);
  int E[NC][8], DP[NC][8];
  {
    int A[NC][8], B[NC][8];
    fetch53<NC, U16>(D, J, 2 * J.jbeg - 2, A);
    fetch53<NC, U16>(D, J, 2 * J.jbeg - 1, B);
    fetch53<NC, U16>(D, J, 2 * J.jbeg, E);
#pragma unroll
    for(int c = 0; c < NC; ++c)
#pragma unroll
      for(int i = 0; i < 8; ++i)
        DP[c][i] = B[c][i] - ((A[c][i] + E[c][i]) >> 1);
  }
  for(int j = J.jbeg; j < J.jend; ++j)
  {
    int O[NC][8], E2[NC][8];
    fetch53<NC, U16>(D, J, 2 * j + 1, O);
    fetch53<NC, U16>(D, J, 2 * j + 2, E2);
#pragma unroll 1
    for(int c = 0; c < NC; ++c)
    {
      int s[8], d[8];
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        d[i] = O[c][i] - ((E[c][i] + E2[c][i]) >> 1);
        s[i] = E[c][i] + ((DP[c][i] + d[i] + 2) >> 2);
        if(J.hn == 1)
        {
          s[i] = E[c][i];
          d[i] = O[c][i] << 1;
        }
        DP[c][i] = d[i];
        E[c][i] = E2[c][i];
      }
      int lo[4], hi[4];
      hfwd53<true>(s, J.wn, lo, hi);
      store_band_rows(D, J, g, c, j, false, lo, hi);
      hfwd53<true>(d, J.wn, lo, hi);
      store_band_rows(D, J, g, c, j, true, lo, hi);
    }
  }
}

/* =============================================================================================
 * forward 5/3
 * =========================================================================================== */
template <int NC, bool U16, int STAGES>
__global__ void __launch_bounds__(B2K_WARPS_PER_CTA * 32) k_dwt53_fwd(const DwtLevelDesc* __restrict__ descs)
{
  extern __shared__ __align__(16) uint8_t smem_dwt[];
  typedef RowStage<NC, U16> RS;
  const DwtLevelDesc D = descs[blockIdx.y]; /* by value: fields live in (uniform) registers, not re-read after every store */
  Job J;
  if(!decode_job(D, J))
    return;
  if(J.hn == 1 || J.wn == 1)
  {
    fwd53_degenerate_job<NC, U16>(descs + blockIdx.y, J);
    return;
  }
  const BandGeom g = band_geom(D);
  const StoreCtx SC = store_ctx(D, J, g);
  uint8_t* wsm = smem_dwt + (size_t)(threadIdx.x >> 5) * STAGES * RS::PAIRB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dwt + (size_t)B2K_WARPS_PER_CTA * STAGES * RS::PAIRB) + (threadIdx.x >> 5) * STAGES;
  const bool fast = RS::lane_fast(D, J);
  const unsigned fastmask = __ballot_sync(0xffffffffu, fast);
  /* interior strip: rows arrive by bulk copy (TMA engine) on per-slot mbarriers; edge strips keep the LDGSTS path */
  const bool bulk = !U16 && fastmask == 0xffffffffu;
  if(bulk)
  {
    if(J.lane == 0)
    {
#pragma unroll
      for(int s = 0; s < STAGES; ++s)
        mbar_init(bars + s, 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  int mcol[8]; /* mirrored column of each of the lane's samples: only edge lanes use them */
#pragma unroll
  for(int i = 0; i < 8; ++i)
    mcol[i] = mirror_rel(J.ulane - D.u0 + i, J.wn);

  /* pairs t = jbeg-1 .. jend-1 : rows (2t+1, 2t+2); the even row before them is fetched directly */
  const int tfirst = J.jbeg - 1, tlast = J.jend - 1;
  int tfill = tfirst;
  auto fill_pair = [&](int tf) {
    const int slot = (tf - tfirst) % STAGES;
    uint8_t* st = wsm + (size_t)slot * RS::PAIRB;
    if(bulk)
    {
      if(J.lane == 0)
      {
        mbar_expect_tx(bars + slot, RS::PAIRB);
        RS::fill_bulk(st, 0, D, J, 2 * tf + 1, bars + slot);
        RS::fill_bulk(st, 1, D, J, 2 * tf + 2, bars + slot);
      }
    }
    else
    {
      RS::fill(st, 0, D, J, 2 * tf + 1, fast, fastmask, mcol);
      RS::fill(st, 1, D, J, 2 * tf + 2, fast, fastmask, mcol);
    }
  };
#pragma unroll
  for(int s = 0; s < STAGES - 1; ++s)
  {
    if(tfill <= tlast)
      fill_pair(tfill);
    cp_async_commit();
    ++tfill;
  }
  int E[NC][8], DP[NC][8];
  fetch53<NC, U16>(D, J, 2 * tfirst, E);
#pragma unroll
  for(int c = 0; c < NC; ++c)
#pragma unroll
    for(int i = 0; i < 8; ++i)
      DP[c][i] = 0;

  for(int t = tfirst; t <= tlast; ++t)
  {
    __syncwarp(); /* every lane has read the stage that is refilled next */
    if(tfill <= tlast)
      fill_pair(tfill);
    cp_async_commit();
    ++tfill;
    const uint8_t* st = wsm + (size_t)((t - tfirst) % STAGES) * RS::PAIRB;
    int O[NC][8], E2[NC][8];
    if(bulk)
    {
      mbar_wait(bars + (t - tfirst) % STAGES, (unsigned)(((t - tfirst) / STAGES) & 1));
      RS::read_linear(st, 0, J, O);
      RS::read_linear(st, 1, J, E2);
    }
    else
    {
      cp_async_wait<STAGES - 1>();
      __syncwarp(); /* rows were copied cooperatively */
      RS::read(st, 0, D, J, O);
      RS::read(st, 1, D, J, E2);
    }
    rct_fwd_inplace<NC>(D, O);
    rct_fwd_inplace<NC>(D, E2);
    const bool emit = t >= J.jbeg;
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      int s[8], d[8];
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        d[i] = O[c][i] - ((E[c][i] + E2[c][i]) >> 1);
        s[i] = E[c][i] + ((DP[c][i] + d[i] + 2) >> 2);
        DP[c][i] = d[i];
        E[c][i] = E2[c][i];
      }
      if(emit)
      { /* warp-uniform */
        int lo[4], hi[4];
        hfwd53<false>(s, J.wn, lo, hi);
        store_rows_fast(D, g, SC, c, t, false, lo, hi);
        hfwd53<false>(d, J.wn, lo, hi);
        store_rows_fast(D, g, SC, c, t, true, lo, hi);
      }
    }
  }
  cp_async_wait<0>();
}

/* =============================================================================================
 * forward 9/7: vertical pipeline  d1[t] -> s1[t] -> d2[t-1] -> s2[t-1]   
 * =========================================================================================== */
/* unpipelined path for lines of one sample (cold) */
template <int NC, bool U16>
__device__ __noinline__ void fwd97_degenerate_job(const DwtLevelDesc* __restrict__ dptr, const Job J)
{
  const DwtLevelDesc& D = *dptr;
  const BandGeom g = band_geom(D);
  const float invK = (float)(1.0 / 1.230174105);
  const float deltaS = __fmul_rn(F97_DELTA, invK);

  float Ev[NC][8], D1[NC][8], S1[NC][8], D2[NC][8];
  fetch97<NC, U16>(D, J, 2 * (J.jbeg - 2), Ev);
#pragma unroll
  for(int c = 0; c < NC; ++c)
#pragma unroll
    for(int i = 0; i < 8; ++i)
      D1[c][i] = S1[c][i] = D2[c][i] = 0.f;

  for(int t = J.jbeg - 2; t <= J.jend; ++t)
  {
    float O[NC][8], E2[NC][8];
    fetch97<NC, U16>(D, J, 2 * t + 1, O);
    fetch97<NC, U16>(D, J, 2 * t + 2, E2);
    const bool emit = (t - 1) >= J.jbeg;
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      float lowrow[8], highrow[8];
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        const float d1 = fmaf(Ev[c][i] + E2[c][i], F97_ALPHA, O[c][i]);
        const float s1 = fmaf(D1[c][i] + d1, F97_BETA, Ev[c][i]);
        const float d2 = __fmul_rn(fmaf(S1[c][i] + s1, F97_GAMMA, D1[c][i]), F97_K);
        const float s2 = __fmul_rn(fmaf(D2[c][i] + d2, deltaS, S1[c][i]), invK);
        lowrow[i] = s2;
        highrow[i] = d2;
        if(J.hn == 1)
        { /* WaveletFwd.cpp L639-654; the values of pair t-1 are asked for, every row mirrors
             to the single real one */
          lowrow[i] = Ev[c][i];
          highrow[i] = __fmul_rn(O[c][i], 2.0f);
        }
        D2[c][i] = d2;
        D1[c][i] = d1;
        S1[c][i] = s1;
        Ev[c][i] = E2[c][i];
      }
      /* shuffles are executed by every lane on every iteration */
      int lo[4], hi[4];
      hfwd97(lowrow, J.wn, invK, deltaS, lo, hi);
      if(emit)
        store_band_rows(D, J, g, c, t - 1, false, lo, hi);
      hfwd97(highrow, J.wn, invK, deltaS, lo, hi);
      if(emit)
        store_band_rows(D, J, g, c, t - 1, true, lo, hi);
    }
  }
}


/* float samples of a staged row: finest level = integers from the image (+DC shift, +ICT,
   mct.cpp L584-636), other levels = float bits */
template <int NC>
__device__ __forceinline__ void ict_fwd_convert(const DwtLevelDesc& D, const int (&raw)[NC][8], float (&out)[NC][8])
{
  if(D.first_level)
  {
    const float a_r = 0.299f, a_g = 0.587f, a_b = 0.114f;
    const float cb = __fdiv_rn(0.5f, __fsub_rn(1.0f, a_b)), cr = __fdiv_rn(0.5f, __fsub_rn(1.0f, a_r));
#pragma unroll
    for(int i = 0; i < 8; ++i)
    {
      if(NC == 3)
      {
        const float r = (float)(raw[0][i] + D.shift[0]), g = (float)(raw[NC > 1 ? 1 : 0][i] + D.shift[1]),
                    b = (float)(raw[NC > 2 ? 2 : 0][i] + D.shift[2]);
        /* the reference build contracts a_r*r + a_g*g + a_b*b into two FMAs (pinned against libgrokj2k, tests/test_interop.py) */
        const float y = __fmaf_rn(a_b, b, __fmaf_rn(a_g, g, __fmul_rn(a_r, r)));
        out[0][i] = y;
        out[NC > 1 ? 1 : 0][i] = __fmul_rn(cb, __fsub_rn(b, y));
        out[NC > 2 ? 2 : 0][i] = __fmul_rn(cr, __fsub_rn(r, y));
      }
      else
        out[0][i] = (float)(raw[0][i] + D.shift[0]);
    }
  }
  else
  {
#pragma unroll
    for(int c = 0; c < NC; ++c)
#pragma unroll
      for(int i = 0; i < 8; ++i)
        out[c][i] = __int_as_float(raw[c][i]);
  }
}

template <int NC, bool U16, int STAGES>
__global__ void __launch_bounds__(B2K_WARPS_PER_CTA * 32) k_dwt97_fwd(const DwtLevelDesc* __restrict__ descs)
{
  extern __shared__ __align__(16) uint8_t smem_dwt[];
  typedef RowStage<NC, U16> RS;
  const DwtLevelDesc D = descs[blockIdx.y]; /* by value */
  Job J;
  if(!decode_job(D, J))
    return;
  if(J.hn == 1 || J.wn == 1)
  {
    fwd97_degenerate_job<NC, U16>(descs + blockIdx.y, J);
    return;
  }
  const BandGeom g = band_geom(D);
  const StoreCtx SC = store_ctx(D, J, g);
  uint8_t* wsm = smem_dwt + (size_t)(threadIdx.x >> 5) * STAGES * RS::PAIRB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dwt + (size_t)B2K_WARPS_PER_CTA * STAGES * RS::PAIRB) + (threadIdx.x >> 5) * STAGES;
  const bool fast = RS::lane_fast(D, J);
  const unsigned fastmask = __ballot_sync(0xffffffffu, fast);
  const bool bulk = !U16 && fastmask == 0xffffffffu; /* interior strip: rows by bulk copy (TMA engine), see k_dwt53_fwd */
  if(bulk)
  {
    if(J.lane == 0)
    {
#pragma unroll
      for(int s = 0; s < STAGES; ++s)
        mbar_init(bars + s, 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  int mcol[8];
#pragma unroll
  for(int i = 0; i < 8; ++i)
    mcol[i] = mirror_rel(J.ulane - D.u0 + i, J.wn);
  const float invK = (float)(1.0 / 1.230174105);
  const float deltaS = __fmul_rn(F97_DELTA, invK);

  /* pairs t = jbeg-2 .. jend : rows (2t+1, 2t+2); output pair t-1 from t = jbeg+1 on */
  const int tfirst = J.jbeg - 2, tlast = J.jend;
  int tfill = tfirst;
  auto fill_pair = [&](int tf) {
    const int slot = (tf - tfirst) % STAGES;
    uint8_t* st = wsm + (size_t)slot * RS::PAIRB;
    if(bulk)
    {
      if(J.lane == 0)
      {
        mbar_expect_tx(bars + slot, RS::PAIRB);
        RS::fill_bulk(st, 0, D, J, 2 * tf + 1, bars + slot);
        RS::fill_bulk(st, 1, D, J, 2 * tf + 2, bars + slot);
      }
    }
    else
    {
      RS::fill(st, 0, D, J, 2 * tf + 1, fast, fastmask, mcol);
      RS::fill(st, 1, D, J, 2 * tf + 2, fast, fastmask, mcol);
    }
  };
#pragma unroll
  for(int s = 0; s < STAGES - 1; ++s)
  {
    if(tfill <= tlast)
      fill_pair(tfill);
    cp_async_commit();
    ++tfill;
  }
  float Ev[NC][8], D1[NC][8], S1[NC][8], D2[NC][8];
  fetch97<NC, U16>(D, J, 2 * tfirst, Ev);
#pragma unroll
  for(int c = 0; c < NC; ++c)
#pragma unroll
    for(int i = 0; i < 8; ++i)
      D1[c][i] = S1[c][i] = D2[c][i] = 0.f;

  for(int t = tfirst; t <= tlast; ++t)
  {
    __syncwarp();
    if(tfill <= tlast)
      fill_pair(tfill);
    cp_async_commit();
    ++tfill;
    if(bulk)
      mbar_wait(bars + (t - tfirst) % STAGES, (unsigned)(((t - tfirst) / STAGES) & 1));
    else
    {
      cp_async_wait<STAGES - 1>();
      __syncwarp();
    }
    const uint8_t* st = wsm + (size_t)((t - tfirst) % STAGES) * RS::PAIRB;
    float O[NC][8], E2[NC][8];
    {
      int raw[NC][8];
      if(bulk)
        RS::read_linear(st, 0, J, raw);
      else
        RS::read(st, 0, D, J, raw);
      ict_fwd_convert<NC>(D, raw, O);
      if(bulk)
        RS::read_linear(st, 1, J, raw);
      else
        RS::read(st, 1, D, J, raw);
      ict_fwd_convert<NC>(D, raw, E2);
    }
    const bool emit = (t - 1) >= J.jbeg;
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      float lowrow[8], highrow[8];
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        const float d1 = fmaf(Ev[c][i] + E2[c][i], F97_ALPHA, O[c][i]);
        const float s1 = fmaf(D1[c][i] + d1, F97_BETA, Ev[c][i]);
        const float d2 = __fmul_rn(fmaf(S1[c][i] + s1, F97_GAMMA, D1[c][i]), F97_K);
        const float s2 = __fmul_rn(fmaf(D2[c][i] + d2, deltaS, S1[c][i]), invK);
        lowrow[i] = s2;
        highrow[i] = d2;
        D2[c][i] = d2;
        D1[c][i] = d1;
        S1[c][i] = s1;
        Ev[c][i] = E2[c][i];
      }
      if(emit)
      { /* warp-uniform */
        int lo[4], hi[4];
        hfwd97(lowrow, 2, invK, deltaS, lo, hi);
        store_rows_fast(D, g, SC, c, t - 1, false, lo, hi);
        hfwd97(highrow, 2, invK, deltaS, lo, hi);
        store_rows_fast(D, g, SC, c, t - 1, true, lo, hi);
      }
    }
  }
  cp_async_wait<0>();
}

/* =============================================================================================
 * inverse: descriptor roles swap -- out_ll / out_c are the SOURCE (LL and HL/LH/HH), in[] the
 * destination (interleaved samples of the next finer resolution, or the image at the finest).
 * =========================================================================================== */
template <int NC>
__device__ __forceinline__ void fetch_band_rows(const DwtLevelDesc& D, const Job& J, const BandGeom& g, int j,
                                                bool vhigh, int (&lo)[NC][4], int (&hi)[NC][4])
{
  /* vertical mirror in the interleaved domain */
  const int vm = D.v0 + mirror_rel(2 * j + (vhigh ? 1 : 0) - D.v0, J.hn);
  const int jm = vm >> 1;
  const int k0 = J.ulane >> 1;
#pragma unroll
  for(int c = 0; c < NC; ++c)
  {
    if(!J.need || (vm & 1) != (vhigh ? 1 : 0))
    {
#pragma unroll
      for(int i = 0; i < 4; ++i)
        lo[c][i] = hi[c][i] = 0;
      continue;
    }
    const int32_t* lrow;
    const int32_t* hrow;
    if(!vhigh)
    {
      lrow = reinterpret_cast<const int32_t*>(D.out_ll[c]) + (size_t)(jm - g.y0l) * D.ll_pitch;
      hrow = reinterpret_cast<const int32_t*>(D.out_c[c]) + (size_t)(jm - g.y0l) * D.c_pitch + g.snx;
    }
    else
    {
      lrow = reinterpret_cast<const int32_t*>(D.out_c[c]) + (size_t)(g.sny + jm - g.y0h) * D.c_pitch;
      hrow = lrow + g.snx;
    }
    load4_band(lrow, k0, g.x0l, D.u0, D.u1, 0, lo[c]);
    load4_band(hrow, k0, g.x0h, D.u0, D.u1, 1, hi[c]);
  }
}

/* inverse horizontal 5/3: WaveletReverse.cpp L879-1072 */
__device__ __forceinline__ void hinv53(const int (&lo)[4], const int (&hi)[4], int wn, int (&r)[8])
{
  const int dm = __shfl_up_sync(0xffffffffu, hi[3], 1);
  int e[5];
  e[0] = lo[0] - ((dm + hi[0] + 2) >> 2);
  e[1] = lo[1] - ((hi[0] + hi[1] + 2) >> 2);
  e[2] = lo[2] - ((hi[1] + hi[2] + 2) >> 2);
  e[3] = lo[3] - ((hi[2] + hi[3] + 2) >> 2);
  e[4] = __shfl_down_sync(0xffffffffu, e[0], 1);
#pragma unroll
  for(int i = 0; i < 4; ++i)
  {
    r[2 * i] = e[i];
    r[2 * i + 1] = hi[i] + ((e[i] + e[i + 1]) >> 1);
  }
  if(wn == 1)
  {
#pragma unroll
    for(int i = 0; i < 4; ++i)
    {
      r[2 * i] = lo[i];
      r[2 * i + 1] = hi[i] >> 1;
    }
  }
}

/* inverse horizontal 9/7: WaveletReverse97.cpp L837-857, constants L98-103 */
__device__ __forceinline__ void hinv97(const int (&loi)[4], const int (&hii)[4], int wn, float (&r)[8])
{
  const float K = 1.230174105f, twice_invK = 1.625732422f;
  float s[5], d[4];
#pragma unroll
  for(int i = 0; i < 4; ++i)
  {
    s[i] = __fmul_rn(__int_as_float(loi[i]), K);
    d[i] = __fmul_rn(__int_as_float(hii[i]), twice_invK);
  }
  float dm = __shfl_up_sync(0xffffffffu, d[3], 1);
  s[0] = fmaf(dm + d[0], -0.443506852f, s[0]);
#pragma unroll
  for(int i = 1; i < 4; ++i)
    s[i] = fmaf(d[i - 1] + d[i], -0.443506852f, s[i]);
  s[4] = __shfl_down_sync(0xffffffffu, s[0], 1);
#pragma unroll
  for(int i = 0; i < 4; ++i)
    d[i] = fmaf(s[i] + s[i + 1], -0.882911075f, d[i]);
  dm = __shfl_up_sync(0xffffffffu, d[3], 1);
  s[0] = fmaf(dm + d[0], 0.052980118f, s[0]);
#pragma unroll
  for(int i = 1; i < 4; ++i)
    s[i] = fmaf(d[i - 1] + d[i], 0.052980118f, s[i]);
  s[4] = __shfl_down_sync(0xffffffffu, s[0], 1);
#pragma unroll
  for(int i = 0; i < 4; ++i)
  {
    r[2 * i] = s[i];
    r[2 * i + 1] = fmaf(s[i] + s[i + 1], 1.586134342f, d[i]);
  }
  if(wn == 1)
  {
#pragma unroll
    for(int i = 0; i < 4; ++i)
    {
      r[2 * i] = __int_as_float(loi[i]);
      r[2 * i + 1] = __int_as_float(hii[i]);
    }
  }
}

/* write one reconstructed sample row (canvas row v) of NC components */
template <int NC>
__device__ __forceinline__ void store_rows53(const DwtLevelDesc& D, const Job& J, int v, int (&x)[NC][8])
{
  if(!J.owner || v < D.v0 || v >= D.v1)
    return;
  unsigned m = 0;
#pragma unroll
  for(int i = 0; i < 8; ++i)
    if(J.ulane + i >= D.u0 && J.ulane + i < D.u1)
      m |= 1u << i;
  if(D.first_level)
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
    {
      if(NC == 3)
      { /* mct.cpp L201-256 */
        const int y = x[0][i], u = x[1][i], w = x[2][i];
        const int gg = y - ((u + w) >> 2);
        x[0][i] = w + gg;
        x[1][i] = gg;
        x[2][i] = u + gg;
      }
#pragma unroll
      for(int c = 0; c < NC; ++c)
        x[c][i] = min(max(x[c][i] - D.shift[c], D.lo[c]), D.hi[c]);
    }
  }
#pragma unroll
  for(int c = 0; c < NC; ++c)
  {
    int32_t* row = reinterpret_cast<int32_t*>(const_cast<void*>(D.in[c])) + (size_t)(v - D.v0) * D.in_pitch;
    store8(row, J.ulane - D.u0, x[c], m);
  }
}
template <int NC>
__device__ __forceinline__ void store_rows97(const DwtLevelDesc& D, const Job& J, int v, float (&x)[NC][8])
{
  if(!J.owner || v < D.v0 || v >= D.v1)
    return;
  unsigned m = 0;
#pragma unroll
  for(int i = 0; i < 8; ++i)
    if(J.ulane + i >= D.u0 && J.ulane + i < D.u1)
      m |= 1u << i;
  int o[NC][8];
  if(D.first_level)
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
    {
      float f[NC];
      if(NC == 3)
      { /* mct.cpp L318-391 */
        const float y = x[0][i], u = x[1][i], w = x[2][i];
        f[0] = __fmaf_rn(w, 1.402f, y); /* FMA / FNMA as the reference build contracts them (tests/test_interop.py) */
        if(NC > 1)
          f[NC > 1 ? 1 : 0] = __fmaf_rn(-w, 0.71414f, __fmaf_rn(-u, 0.34413f, y));
        if(NC > 2)
          f[NC > 2 ? 2 : 0] = __fmaf_rn(u, 1.772f, y);
      }
      else
        f[0] = x[0][i];
#pragma unroll
      for(int c = 0; c < NC; ++c)
        o[c][i] = min(max(__float2int_rn(f[c]) - D.shift[c], D.lo[c]), D.hi[c]);
    }
  }
  else
  {
#pragma unroll
    for(int c = 0; c < NC; ++c)
#pragma unroll
      for(int i = 0; i < 8; ++i)
        o[c][i] = __float_as_int(x[c][i]);
  }
#pragma unroll
  for(int c = 0; c < NC; ++c)
  {
    int32_t* row = reinterpret_cast<int32_t*>(const_cast<void*>(D.in[c])) + (size_t)(v - D.v0) * D.in_pitch;
    store8(row, J.ulane - D.u0, o[c], m);
  }
}

/* a resolution one sample wide or high: straightforward, unpipelined path (cold) */
template <int NC>
__device__ __noinline__ void inv53_degenerate_job(const DwtLevelDesc* __restrict__ dptr, const Job J)
{
  const DwtLevelDesc& D = *dptr;
  const BandGeom g = band_geom(D);
  int DV[NC][8], EP[NC][8];
#pragma unroll
  for(int c = 0; c < NC; ++c)
#pragma unroll
    for(int i = 0; i < 8; ++i)
      DV[c][i] = EP[c][i] = 0;

  for(int t = J.jbeg - 1; t <= J.jend; ++t)
  {
    int lo[NC][4], hi[NC][4];
    int sv[NC][8], dv[NC][8];
    fetch_band_rows<NC>(D, J, g, t, false, lo, hi);
#pragma unroll
    for(int c = 0; c < NC; ++c)
      hinv53(lo[c], hi[c], J.wn, sv[c]);
    fetch_band_rows<NC>(D, J, g, t, true, lo, hi);
#pragma unroll
    for(int c = 0; c < NC; ++c)
      hinv53(lo[c], hi[c], J.wn, dv[c]);
    int Er[NC][8], Or[NC][8];
#pragma unroll
    for(int c = 0; c < NC; ++c)
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        const int e = sv[c][i] - ((DV[c][i] + dv[c][i] + 2) >> 2);
        Or[c][i] = DV[c][i] + ((EP[c][i] + e) >> 1);
        Er[c][i] = EP[c][i];
        EP[c][i] = e;
        DV[c][i] = dv[c][i];
      }
    if(J.hn == 1)
    { /* single row: pair t holds it (low unchanged, lone high halved) */
      if(t >= J.jbeg && t < J.jend)
      {
        int hv[NC][8];
#pragma unroll
        for(int c = 0; c < NC; ++c)
#pragma unroll
          for(int i = 0; i < 8; ++i)
            hv[c][i] = dv[c][i] >> 1;
        store_rows53<NC>(D, J, 2 * t, sv);
        store_rows53<NC>(D, J, 2 * t + 1, hv);
      }
    }
    else if(t - 1 >= J.jbeg)
    {
      store_rows53<NC>(D, J, 2 * (t - 1), Er);
      store_rows53<NC>(D, J, 2 * (t - 1) + 1, Or);
    }
  }
}


/* staged band-row fetch for the inverse kernels: per row pair four band rows (LL|HL, LH|HH) x NC,
   each lane owning 4 consecutive samples (16 bytes) of each -> one cp.async per lane per band row,
   512 contiguous bytes per instruction, private slots (no barrier) */
template <int NC>
struct BandStage
{
  static constexpr int ROWB = 512;
  static constexpr int PAIRB = 4 * NC * ROWB;
  struct Lane
  {
    bool need;
    bool fast[4];   /* LL, HL, LH, HH source: 4 samples interior and 16-byte aligned */
    int col[4];     /* band-relative column of the lane's first sample (low, high) per source */
    int mlo[4], mhi[4]; /* mirrored band-relative columns for edge lanes */
  };
  static __device__ __forceinline__ void setup(const DwtLevelDesc& D, const Job& J, const BandGeom& g, Lane& L)
  {
    const int k0 = J.ulane >> 1;
    L.need = J.need;
#pragma unroll
    for(int i = 0; i < 4; ++i)
    {
      const int ul = D.u0 + mirror_rel(2 * (k0 + i) - D.u0, J.wn), uh = D.u0 + mirror_rel(2 * (k0 + i) + 1 - D.u0, J.wn);
      L.mlo[i] = (ul >> 1) - g.x0l;
      L.mhi[i] = (uh >> 1) - g.x0h;
    }
    const bool in_lo = 2 * k0 >= D.u0 && 2 * k0 + 6 < D.u1, in_hi = 2 * k0 + 1 >= D.u0 && 2 * k0 + 7 < D.u1;
    const int clo = k0 - g.x0l, chi = k0 - g.x0h;
    L.col[0] = clo; L.col[1] = g.snx + chi; L.col[2] = clo; L.col[3] = g.snx + chi;
    const bool pitch_ok = ((D.ll_pitch | D.c_pitch) & 3u) == 0;
    L.fast[0] = J.need && pitch_ok && in_lo && (((reinterpret_cast<uintptr_t>(D.out_ll[0]) >> 2) + (unsigned)clo) & 3u) == 0;
    L.fast[1] = J.need && pitch_ok && in_hi && (((reinterpret_cast<uintptr_t>(D.out_c[0]) >> 2) + (unsigned)(g.snx + chi)) & 3u) == 0;
    L.fast[2] = J.need && pitch_ok && in_lo && (((reinterpret_cast<uintptr_t>(D.out_c[0]) >> 2) + (unsigned)clo) & 3u) == 0;
    L.fast[3] = L.fast[1];
  }
  /* band rows of pair t into the stage */
  static __device__ __forceinline__ void fill(uint8_t* stage, const DwtLevelDesc& D, const Job& J, const BandGeom& g,
                                              const Lane& L, int t)
  {
    if(!L.need)
      return;
    const int jl = ((D.v0 + mirror_rel(2 * t - D.v0, J.hn)) >> 1), jh = ((D.v0 + mirror_rel(2 * t + 1 - D.v0, J.hn)) >> 1);
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      const int32_t* src[4];
      src[0] = reinterpret_cast<const int32_t*>(D.out_ll[c]) + (jl - g.y0l) * (int)D.ll_pitch;
      src[1] = reinterpret_cast<const int32_t*>(D.out_c[c]) + (jl - g.y0l) * (int)D.c_pitch;
      src[2] = reinterpret_cast<const int32_t*>(D.out_c[c]) + (g.sny + jh - g.y0h) * (int)D.c_pitch;
      src[3] = src[2];
#pragma unroll
      for(int b = 0; b < 4; ++b)
      {
        uint8_t* dst = stage + (b * NC + c) * ROWB + J.lane * 16;
        if(L.fast[b])
          cp_async16(dst, src[b] + L.col[b]);
        else
        {
          const int base = (b & 1) ? g.snx : 0;
#pragma unroll
          for(int i = 0; i < 4; ++i)
            cp_async4(dst + 4 * i, src[b] + base + ((b & 1) ? L.mhi[i] : L.mlo[i]));
        }
      }
    }
  }
  /* bulk path: every lane of the warp is `fast` on all four sources, so each band row of the pair is 512 contiguous,
     16-byte aligned bytes starting at lane 0's column: lane 0 asks the copy engine (TMA, UBLKCP) for the 4 x NC rows,
     which land in the same linear layout the per-lane cp.async path writes and complete on `bar` */
  static __device__ __forceinline__ void fill_bulk(uint8_t* stage, const DwtLevelDesc& D, const Job& J, const BandGeom& g,
                                                   const Lane& L, int t, uint64_t* bar)
  {
    const int jl = ((D.v0 + mirror_rel(2 * t - D.v0, J.hn)) >> 1), jh = ((D.v0 + mirror_rel(2 * t + 1 - D.v0, J.hn)) >> 1);
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      const int32_t* src[4];
      src[0] = reinterpret_cast<const int32_t*>(D.out_ll[c]) + (jl - g.y0l) * (int)D.ll_pitch;
      src[1] = reinterpret_cast<const int32_t*>(D.out_c[c]) + (jl - g.y0l) * (int)D.c_pitch;
      src[2] = reinterpret_cast<const int32_t*>(D.out_c[c]) + (g.sny + jh - g.y0h) * (int)D.c_pitch;
      src[3] = src[2];
#pragma unroll
      for(int b = 0; b < 4; ++b)
        bulk_g2s(stage + (b * NC + c) * ROWB, src[b] + L.col[b], ROWB, bar);
    }
  }
  static __device__ __forceinline__ bool all_fast(const Lane& L)
  {
    return __all_sync(0xffffffffu, L.fast[0] && L.fast[1] && L.fast[2] && L.fast[3]);
  }
  static __device__ __forceinline__ void read(const uint8_t* stage, const Job& J, int b, int c, int (&v)[4])
  {
    const int4 a = *reinterpret_cast<const int4*>(stage + (b * NC + c) * ROWB + J.lane * 16);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
};

/* per-lane constants of the reconstructed-row stores */
struct OutCtx
{
  unsigned m;
  bool vec;
  int col;
};
template <bool OUT16 = false>
__device__ __forceinline__ OutCtx out_ctx(const DwtLevelDesc& D, const Job& J)
{
  OutCtx o;
  o.m = 0;
  if(J.owner)
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
      if(J.ulane + i >= D.u0 && J.ulane + i < D.u1)
        o.m |= 1u << i;
  }
  o.col = J.ulane - D.u0;
  if(OUT16)
    o.vec = o.m == 0xFF && (D.in_pitch & 7u) == 0 && (((reinterpret_cast<uintptr_t>(D.in[0]) >> 1) + (unsigned)o.col) & 7u) == 0;
  else
    o.vec = o.m == 0xFF && (D.in_pitch & 3u) == 0 && (((reinterpret_cast<uintptr_t>(D.in[0]) >> 2) + (unsigned)o.col) & 3u) == 0;
  return o;
}
template <int NC, bool OUT16 = false>
__device__ __forceinline__ void store_rows53_fast(const DwtLevelDesc& D, const OutCtx& O, int v, int (&x)[NC][8])
{
  if(v < D.v0 || v >= D.v1 || O.m == 0)
    return;
  if(D.first_level)
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
    {
      if(NC == 3)
      { /* mct.cpp L201-256 */
        const int y = x[0][i], u = x[NC > 1 ? 1 : 0][i], w = x[NC > 2 ? 2 : 0][i];
        const int gg = y - ((u + w) >> 2);
        x[0][i] = w + gg;
        x[NC > 1 ? 1 : 0][i] = gg;
        x[NC > 2 ? 2 : 0][i] = u + gg;
      }
#pragma unroll
      for(int c = 0; c < NC; ++c)
        x[c][i] = min(max(x[c][i] - D.shift[c], D.lo[c]), D.hi[c]);
    }
  }
#pragma unroll
  for(int c = 0; c < NC; ++c)
  {
    if(OUT16)
    { /* 16-bit sample containers: the clamped value fits, two's complement for signed data */
      uint16_t* q = reinterpret_cast<uint16_t*>(const_cast<void*>(D.in[c])) + ((v - D.v0) * (int)D.in_pitch + O.col);
      if(O.vec)
        *reinterpret_cast<uint4*>(q) = make_uint4((x[c][0] & 0xFFFF) | (x[c][1] << 16), (x[c][2] & 0xFFFF) | (x[c][3] << 16),
                                                  (x[c][4] & 0xFFFF) | (x[c][5] << 16), (x[c][6] & 0xFFFF) | (x[c][7] << 16));
      else
      {
#pragma unroll
        for(int i = 0; i < 8; ++i)
          if(O.m & (1u << i))
            q[i] = (uint16_t)x[c][i];
      }
      continue;
    }
    int32_t* p = reinterpret_cast<int32_t*>(const_cast<void*>(D.in[c])) + ((v - D.v0) * (int)D.in_pitch + O.col);
    if(O.vec)
    {
      reinterpret_cast<int4*>(p)[0] = make_int4(x[c][0], x[c][1], x[c][2], x[c][3]);
      reinterpret_cast<int4*>(p)[1] = make_int4(x[c][4], x[c][5], x[c][6], x[c][7]);
    }
    else
    {
#pragma unroll
      for(int i = 0; i < 8; ++i)
        if(O.m & (1u << i))
          p[i] = x[c][i];
    }
  }
}

template <bool DEGEN = true>
__device__ __forceinline__ void hinv53t(const int (&lo)[4], const int (&hi)[4], int (&r)[8])
{
  const int dm = __shfl_up_sync(0xffffffffu, hi[3], 1);
  int e[5];
  e[0] = lo[0] - ((dm + hi[0] + 2) >> 2);
  e[1] = lo[1] - ((hi[0] + hi[1] + 2) >> 2);
  e[2] = lo[2] - ((hi[1] + hi[2] + 2) >> 2);
  e[3] = lo[3] - ((hi[2] + hi[3] + 2) >> 2);
  e[4] = __shfl_down_sync(0xffffffffu, e[0], 1);
#pragma unroll
  for(int i = 0; i < 4; ++i)
  {
    r[2 * i] = e[i];
    r[2 * i + 1] = hi[i] + ((e[i] + e[i + 1]) >> 1);
  }
}

template <int NC, int STAGES, bool OUT16>
__global__ void __launch_bounds__(B2K_WARPS_PER_CTA * 32) k_dwt53_inv(const DwtLevelDesc* __restrict__ descs)
{
  extern __shared__ __align__(16) uint8_t smem_dwt[];
  typedef BandStage<NC> BS;
  const DwtLevelDesc D = descs[blockIdx.y]; /* by value: fields live in (uniform) registers */
  Job J;
  if(!decode_job(D, J))
    return;
  if((J.hn == 1 || J.wn == 1) && !OUT16)
  {
    inv53_degenerate_job<NC>(descs + blockIdx.y, J);
    return;
  }
  const BandGeom g = band_geom(D);
  typename BS::Lane L;
  BS::setup(D, J, g, L);
  const OutCtx OC = out_ctx<OUT16>(D, J);
  uint8_t* wsm = smem_dwt + (size_t)(threadIdx.x >> 5) * STAGES * BS::PAIRB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dwt + (size_t)B2K_WARPS_PER_CTA * STAGES * BS::PAIRB) + (threadIdx.x >> 5) * STAGES;
  const bool bulk = BS::all_fast(L); /* interior strip: band rows by bulk copy (TMA engine) on per-slot mbarriers */
  if(bulk)
  {
    if(J.lane == 0)
    {
#pragma unroll
      for(int s = 0; s < STAGES; ++s)
        mbar_init(bars + s, 1);
      mbar_fence_init();
    }
    __syncwarp();
  }

  const int tfirst = J.jbeg - 1, tlast = J.jend;
  auto fill_pair = [&](int tf) {
    const int slot = (tf - tfirst) % STAGES;
    uint8_t* st = wsm + (size_t)slot * BS::PAIRB;
    if(bulk)
    {
      if(J.lane == 0)
      {
        mbar_expect_tx(bars + slot, BS::PAIRB);
        BS::fill_bulk(st, D, J, g, L, tf, bars + slot);
      }
    }
    else
      BS::fill(st, D, J, g, L, tf);
  };
  int tfill = tfirst;
#pragma unroll
  for(int s = 0; s < STAGES - 1; ++s)
  {
    if(tfill <= tlast)
      fill_pair(tfill);
    cp_async_commit();
    ++tfill;
  }
  int DV[NC][8], EP[NC][8];
#pragma unroll
  for(int c = 0; c < NC; ++c)
#pragma unroll
    for(int i = 0; i < 8; ++i)
      DV[c][i] = EP[c][i] = 0;

  for(int t = tfirst; t <= tlast; ++t)
  {
    if(bulk)
      __syncwarp(); /* every lane has read the slot that is refilled next */
    if(tfill <= tlast)
      fill_pair(tfill);
    cp_async_commit();
    ++tfill;
    if(bulk)
      mbar_wait(bars + (t - tfirst) % STAGES, (unsigned)(((t - tfirst) / STAGES) & 1));
    else
      cp_async_wait<STAGES - 1>();
    const uint8_t* st = wsm + (size_t)((t - tfirst) % STAGES) * BS::PAIRB;
    int Er[NC][8], Or[NC][8];
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      int lo[4], hi[4], sv[8], dv[8];
      BS::read(st, J, 0, c, lo);
      BS::read(st, J, 1, c, hi);
      hinv53t<false>(lo, hi, sv);
      BS::read(st, J, 2, c, lo);
      BS::read(st, J, 3, c, hi);
      hinv53t<false>(lo, hi, dv);
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        const int e = sv[i] - ((DV[c][i] + dv[i] + 2) >> 2);
        Or[c][i] = DV[c][i] + ((EP[c][i] + e) >> 1);
        Er[c][i] = EP[c][i];
        EP[c][i] = e;
        DV[c][i] = dv[i];
      }
    }
    if(t - 1 >= J.jbeg)
    {
      store_rows53_fast<NC, OUT16>(D, OC, 2 * (t - 1), Er);
      store_rows53_fast<NC, OUT16>(D, OC, 2 * (t - 1) + 1, Or);
    }
  }
  cp_async_wait<0>();
}

/* unpipelined path for lines of one sample (cold) */
template <int NC>
__device__ __noinline__ void inv97_degenerate_job(const DwtLevelDesc* __restrict__ dptr, const Job J)
{
  const DwtLevelDesc& D = *dptr;
  const BandGeom g = band_geom(D);
  const float K = 1.230174105f, twice_invK = 1.625732422f;
  /* state: d0[t-1], s1[t-1], d1[t-2], s2[t-2] */
  float D0[NC][8], S1[NC][8], D1[NC][8], S2[NC][8];
#pragma unroll
  for(int c = 0; c < NC; ++c)
#pragma unroll
    for(int i = 0; i < 8; ++i)
      D0[c][i] = S1[c][i] = D1[c][i] = S2[c][i] = 0.f;

  for(int t = J.jbeg - 2; t <= J.jend + 1; ++t)
  {
    int lo[NC][4], hi[NC][4];
    float sv[NC][8], dv[NC][8];
    fetch_band_rows<NC>(D, J, g, t, false, lo, hi);
#pragma unroll
    for(int c = 0; c < NC; ++c)
      hinv97(lo[c], hi[c], J.wn, sv[c]);
    fetch_band_rows<NC>(D, J, g, t, true, lo, hi);
#pragma unroll
    for(int c = 0; c < NC; ++c)
      hinv97(lo[c], hi[c], J.wn, dv[c]);
    float Er[NC][8], Or[NC][8];
#pragma unroll
    for(int c = 0; c < NC; ++c)
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        const float s0 = __fmul_rn(sv[c][i], K), d0 = __fmul_rn(dv[c][i], twice_invK);
        const float s1 = fmaf(D0[c][i] + d0, -0.443506852f, s0);            /* s1[t]   */
        const float d1 = fmaf(S1[c][i] + s1, -0.882911075f, D0[c][i]);      /* d1[t-1] */
        const float s2 = fmaf(D1[c][i] + d1, 0.052980118f, S1[c][i]);       /* s2[t-1] */
        const float d2 = fmaf(S2[c][i] + s2, 1.586134342f, D1[c][i]);       /* d2[t-2] */
        Er[c][i] = S2[c][i];
        Or[c][i] = d2;
        D0[c][i] = d0;
        S1[c][i] = s1;
        D1[c][i] = d1;
        S2[c][i] = s2;
      }
    if(J.hn == 1)
    {
      if(t >= J.jbeg && t < J.jend)
      {
        store_rows97<NC>(D, J, 2 * t, sv);
        store_rows97<NC>(D, J, 2 * t + 1, dv);
      }
    }
    else if(t - 2 >= J.jbeg)
    {
      store_rows97<NC>(D, J, 2 * (t - 2), Er);
      store_rows97<NC>(D, J, 2 * (t - 2) + 1, Or);
    }
  }
}


template <int NC>
__device__ __forceinline__ void store_rows97_fast(const DwtLevelDesc& D, const OutCtx& O, int v, float (&x)[NC][8])
{
  if(v < D.v0 || v >= D.v1 || O.m == 0)
    return;
  int o[NC][8];
  if(D.first_level)
  {
#pragma unroll
    for(int i = 0; i < 8; ++i)
    {
      float f[NC];
      if(NC == 3)
      { /* mct.cpp L318-391 */
        const float y = x[0][i], u = x[NC > 1 ? 1 : 0][i], w = x[NC > 2 ? 2 : 0][i];
        f[0] = __fmaf_rn(w, 1.402f, y); /* FMA / FNMA as the reference build contracts them (tests/test_interop.py) */
        f[NC > 1 ? 1 : 0] = __fmaf_rn(-w, 0.71414f, __fmaf_rn(-u, 0.34413f, y));
        f[NC > 2 ? 2 : 0] = __fmaf_rn(u, 1.772f, y);
      }
      else
        f[0] = x[0][i];
#pragma unroll
      for(int c = 0; c < NC; ++c)
        o[c][i] = min(max(__float2int_rn(f[c]) - D.shift[c], D.lo[c]), D.hi[c]);
    }
  }
  else
  {
#pragma unroll
    for(int c = 0; c < NC; ++c)
#pragma unroll
      for(int i = 0; i < 8; ++i)
        o[c][i] = __float_as_int(x[c][i]);
  }
#pragma unroll
  for(int c = 0; c < NC; ++c)
  {
    int32_t* p = reinterpret_cast<int32_t*>(const_cast<void*>(D.in[c])) + ((v - D.v0) * (int)D.in_pitch + O.col);
    if(O.vec)
    {
      reinterpret_cast<int4*>(p)[0] = make_int4(o[c][0], o[c][1], o[c][2], o[c][3]);
      reinterpret_cast<int4*>(p)[1] = make_int4(o[c][4], o[c][5], o[c][6], o[c][7]);
    }
    else
    {
#pragma unroll
      for(int i = 0; i < 8; ++i)
        if(O.m & (1u << i))
          p[i] = o[c][i];
    }
  }
}

template <int NC, int STAGES>
__global__ void __launch_bounds__(B2K_WARPS_PER_CTA * 32) k_dwt97_inv(const DwtLevelDesc* __restrict__ descs)
{
  extern __shared__ __align__(16) uint8_t smem_dwt[];
  typedef BandStage<NC> BS;
  const DwtLevelDesc D = descs[blockIdx.y]; /* by value */
  Job J;
  if(!decode_job(D, J))
    return;
  if(J.hn == 1 || J.wn == 1)
  {
    inv97_degenerate_job<NC>(descs + blockIdx.y, J);
    return;
  }
  const BandGeom g = band_geom(D);
  typename BS::Lane L;
  BS::setup(D, J, g, L);
  const OutCtx OC = out_ctx<false>(D, J);
  uint8_t* wsm = smem_dwt + (size_t)(threadIdx.x >> 5) * STAGES * BS::PAIRB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dwt + (size_t)B2K_WARPS_PER_CTA * STAGES * BS::PAIRB) + (threadIdx.x >> 5) * STAGES;
  const bool bulk = BS::all_fast(L); /* interior strip: band rows by bulk copy (TMA engine) on per-slot mbarriers */
  if(bulk)
  {
    if(J.lane == 0)
    {
#pragma unroll
      for(int s = 0; s < STAGES; ++s)
        mbar_init(bars + s, 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  const float K = 1.230174105f, twice_invK = 1.625732422f;

  const int tfirst = J.jbeg - 2, tlast = J.jend + 1;
  auto fill_pair = [&](int tf) {
    const int slot = (tf - tfirst) % STAGES;
    uint8_t* st = wsm + (size_t)slot * BS::PAIRB;
    if(bulk)
    {
      if(J.lane == 0)
      {
        mbar_expect_tx(bars + slot, BS::PAIRB);
        BS::fill_bulk(st, D, J, g, L, tf, bars + slot);
      }
    }
    else
      BS::fill(st, D, J, g, L, tf);
  };
  int tfill = tfirst;
#pragma unroll
  for(int s = 0; s < STAGES - 1; ++s)
  {
    if(tfill <= tlast)
      fill_pair(tfill);
    cp_async_commit();
    ++tfill;
  }
  /* state: d0[t-1], s1[t-1], d1[t-2], s2[t-2] */
  float D0[NC][8], S1[NC][8], D1[NC][8], S2[NC][8];
#pragma unroll
  for(int c = 0; c < NC; ++c)
#pragma unroll
    for(int i = 0; i < 8; ++i)
      D0[c][i] = S1[c][i] = D1[c][i] = S2[c][i] = 0.f;

  for(int t = tfirst; t <= tlast; ++t)
  {
    if(bulk)
      __syncwarp(); /* every lane has read the slot that is refilled next */
    if(tfill <= tlast)
      fill_pair(tfill);
    cp_async_commit();
    ++tfill;
    if(bulk)
      mbar_wait(bars + (t - tfirst) % STAGES, (unsigned)(((t - tfirst) / STAGES) & 1));
    else
      cp_async_wait<STAGES - 1>();
    const uint8_t* st = wsm + (size_t)((t - tfirst) % STAGES) * BS::PAIRB;
    float Er[NC][8], Or[NC][8];
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      int lo[4], hi[4];
      float sv[8], dv[8];
      BS::read(st, J, 0, c, lo);
      BS::read(st, J, 1, c, hi);
      hinv97(lo, hi, 2, sv);
      BS::read(st, J, 2, c, lo);
      BS::read(st, J, 3, c, hi);
      hinv97(lo, hi, 2, dv);
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        const float s0 = __fmul_rn(sv[i], K), d0 = __fmul_rn(dv[i], twice_invK);
        const float s1 = fmaf(D0[c][i] + d0, -0.443506852f, s0);
        const float d1 = fmaf(S1[c][i] + s1, -0.882911075f, D0[c][i]);
        const float s2 = fmaf(D1[c][i] + d1, 0.052980118f, S1[c][i]);
        const float d2 = fmaf(S2[c][i] + s2, 1.586134342f, D1[c][i]);
        Er[c][i] = S2[c][i];
        Or[c][i] = d2;
        D0[c][i] = d0;
        S1[c][i] = s1;
        D1[c][i] = d1;
        S2[c][i] = s2;
      }
    }
    if(t - 2 >= J.jbeg)
    {
      store_rows97_fast<NC>(D, OC, 2 * (t - 2), Er);
      store_rows97_fast<NC>(D, OC, 2 * (t - 2) + 1, Or);
    }
  }
  cp_async_wait<0>();
}

/* 16-bit sample containers <-> the engine's 32-bit planes: one rectangle (a merged tile row) per
   launch, 8 samples per thread, 128-bit accesses.  Costs 6 B/sample of HBM traffic -- noise next to
   the PCIe transfer it halves. */
__global__ void k_widen16(const uint16_t* __restrict__ src, uint32_t spitch, int32_t* __restrict__ dst, uint32_t dpitch,
                          uint32_t w, uint32_t h, int sgnd)
{
  const uint32_t x8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8, y = blockIdx.y;
  if(x8 >= w || y >= h)
    return;
  const uint16_t* s = src + (size_t)y * spitch + x8;
  int32_t* d = dst + (size_t)y * dpitch + x8;
  if(x8 + 8 <= w && ((reinterpret_cast<uintptr_t>(s) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d) & 15) == 0))
  {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(s));
    int v[8] = {(int)(a.x & 0xFFFF), (int)(a.x >> 16), (int)(a.y & 0xFFFF), (int)(a.y >> 16),
                (int)(a.z & 0xFFFF), (int)(a.z >> 16), (int)(a.w & 0xFFFF), (int)(a.w >> 16)};
    if(sgnd)
    {
#pragma unroll
      for(int i = 0; i < 8; ++i)
        v[i] = (int)(int16_t)v[i];
    }
    reinterpret_cast<int4*>(d)[0] = make_int4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<int4*>(d)[1] = make_int4(v[4], v[5], v[6], v[7]);
  }
  else
    for(uint32_t i = 0; i < 8 && x8 + i < w; ++i)
      d[i] = sgnd ? (int)(int16_t)s[i] : (int)s[i];
}
struct Ptr4
{
  int32_t* p[4];
};
/* pixel-interleaved 16-bit samples (RGB48LE rows, the packed frames of gpup_batch_memory_submit: grok.cpp L1806-1836)
   -> NC int32 planes.  A thread takes 8 pixels: NC 128-bit loads, 2*NC 128-bit stores. */
template <int NC>
__global__ void k_widen16_interleaved(const uint16_t* __restrict__ src, uint32_t spitch, Ptr4 dst, uint32_t dpitch, uint32_t w,
                                      uint32_t h, int sgnd)
{
  const uint32_t x8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8, y = blockIdx.y;
  if(x8 >= w || y >= h)
    return;
  const uint16_t* s = src + (size_t)y * spitch + (size_t)x8 * NC;
  const size_t doff = (size_t)y * dpitch + x8;
  if(x8 + 8 <= w && ((reinterpret_cast<uintptr_t>(s) & 15) == 0) && ((doff & 3) == 0))
  {
    uint32_t wd[4 * NC];
#pragma unroll
    for(int i = 0; i < NC; ++i)
    {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(s) + i);
      wd[4 * i] = a.x; wd[4 * i + 1] = a.y; wd[4 * i + 2] = a.z; wd[4 * i + 3] = a.w;
    }
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
      int v[8];
#pragma unroll
      for(int i = 0; i < 8; ++i)
      {
        const int e = i * NC + c; /* 16-bit element index within the 8-pixel group */
        const uint32_t u = (wd[e >> 1] >> ((e & 1) * 16)) & 0xFFFF;
        v[i] = sgnd ? (int)(int16_t)u : (int)u;
      }
      int4* d = reinterpret_cast<int4*>(dst.p[c] + doff);
      d[0] = make_int4(v[0], v[1], v[2], v[3]);
      d[1] = make_int4(v[4], v[5], v[6], v[7]);
    }
  }
  else
    for(uint32_t i = 0; i < 8 && x8 + i < w; ++i)
#pragma unroll
      for(int c = 0; c < NC; ++c)
      {
        const uint16_t u = s[i * NC + c];
        dst.p[c][doff + i] = sgnd ? (int)(int16_t)u : (int)u;
      }
}
__global__ void k_narrow16(const int32_t* __restrict__ src, uint32_t spitch, uint16_t* __restrict__ dst, uint32_t dpitch,
                           uint32_t w, uint32_t h)
{
  const uint32_t x8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8, y = blockIdx.y;
  if(x8 >= w || y >= h)
    return;
  const int32_t* s = src + (size_t)y * spitch + x8;
  uint16_t* d = dst + (size_t)y * dpitch + x8;
  if(x8 + 8 <= w && ((reinterpret_cast<uintptr_t>(s) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d) & 15) == 0))
  {
    const int4 a = __ldg(reinterpret_cast<const int4*>(s)), b = __ldg(reinterpret_cast<const int4*>(s) + 1);
    *reinterpret_cast<uint4*>(d) = make_uint4((a.x & 0xFFFF) | (a.y << 16), (a.z & 0xFFFF) | (a.w << 16),
                                              (b.x & 0xFFFF) | (b.y << 16), (b.z & 0xFFFF) | (b.w << 16));
  }
  else
    for(uint32_t i = 0; i < 8 && x8 + i < w; ++i)
      d[i] = (uint16_t)s[i];
}

} /* namespace */

void b2k_launch_widen16(const uint16_t* src, uint32_t spitch, int32_t* dst, uint32_t dpitch, uint32_t w, uint32_t h, int sgnd,
                        cudaStream_t st)
{
  if(!w || !h)
    return;
  dim3 grid((w + 8 * 128 - 1) / (8 * 128), h), block(128);
  k_widen16<<<grid, block, 0, st>>>(src, spitch, dst, dpitch, w, h, sgnd);
  b2k_count_launch();
}
void b2k_launch_widen16_interleaved(const uint16_t* src, uint32_t spitch, int32_t* const* dst, int nc, uint32_t dpitch, uint32_t w,
                                    uint32_t h, int sgnd, cudaStream_t st)
{
  if(!w || !h)
    return;
  Ptr4 P{};
  for(int c = 0; c < nc && c < 4; ++c)
    P.p[c] = dst[c];
  dim3 grid((w + 8 * 128 - 1) / (8 * 128), h), block(128);
  switch(nc)
  {
    case 1: k_widen16_interleaved<1><<<grid, block, 0, st>>>(src, spitch, P, dpitch, w, h, sgnd); break;
    case 2: k_widen16_interleaved<2><<<grid, block, 0, st>>>(src, spitch, P, dpitch, w, h, sgnd); break;
    case 3: k_widen16_interleaved<3><<<grid, block, 0, st>>>(src, spitch, P, dpitch, w, h, sgnd); break;
    default: k_widen16_interleaved<4><<<grid, block, 0, st>>>(src, spitch, P, dpitch, w, h, sgnd); break;
  }
  b2k_count_launch();
}
void b2k_launch_narrow16(const int32_t* src, uint32_t spitch, uint16_t* dst, uint32_t dpitch, uint32_t w, uint32_t h,
                         cudaStream_t st)
{
  if(!w || !h)
    return;
  dim3 grid((w + 8 * 128 - 1) / (8 * 128), h), block(128);
  k_narrow16<<<grid, block, 0, st>>>(src, spitch, dst, dpitch, w, h);
  b2k_count_launch();
}

constexpr int FWD_STAGES = 3;
template <int NC, bool U16, int STAGES>
static void launch_fwd53(dim3 grid, dim3 block, cudaStream_t st, const DwtLevelDesc* d)
{
  const size_t smem = (size_t)B2K_WARPS_PER_CTA * STAGES * RowStage<NC, U16>::PAIRB + (size_t)B2K_WARPS_PER_CTA * STAGES * sizeof(uint64_t);
  static DeviceOnce once; /* function attributes are per device */
  once.run([&] {
    cudaFuncSetAttribute(k_dwt53_fwd<NC, U16, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  k_dwt53_fwd<NC, U16, STAGES><<<grid, block, smem, st>>>(d);
}

template <int NC, bool U16, int STAGES>
static void launch_fwd97(dim3 grid, dim3 block, cudaStream_t st, const DwtLevelDesc* d)
{
  const size_t smem = (size_t)B2K_WARPS_PER_CTA * STAGES * RowStage<NC, U16>::PAIRB + (size_t)B2K_WARPS_PER_CTA * STAGES * sizeof(uint64_t);
  static DeviceOnce once; /* function attributes are per device */
  once.run([&] {
    cudaFuncSetAttribute(k_dwt97_fwd<NC, U16, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  k_dwt97_fwd<NC, U16, STAGES><<<grid, block, smem, st>>>(d);
}

void b2k_launch_dwt_fwd(const DwtLevelDesc* d, int ndesc, int max_jobs, int nc, bool irreversible, bool in_u16,
                        cudaStream_t st)
{
  if(ndesc <= 0 || max_jobs <= 0)
    return;
  dim3 grid((max_jobs + B2K_WARPS_PER_CTA - 1) / B2K_WARPS_PER_CTA, ndesc), block(B2K_WARPS_PER_CTA * 32);
  if(!irreversible)
  {
    if(nc == 3)
    {
      if(in_u16) launch_fwd53<3, true, FWD_STAGES>(grid, block, st, d);
      else launch_fwd53<3, false, FWD_STAGES>(grid, block, st, d);
    }
    else
    {
      if(in_u16) launch_fwd53<1, true, FWD_STAGES>(grid, block, st, d);
      else launch_fwd53<1, false, FWD_STAGES>(grid, block, st, d);
    }
  }
  else
  {
    if(nc == 3)
    {
      if(in_u16) launch_fwd97<3, true, FWD_STAGES>(grid, block, st, d);
      else launch_fwd97<3, false, FWD_STAGES>(grid, block, st, d);
    }
    else
    {
      if(in_u16) launch_fwd97<1, true, FWD_STAGES>(grid, block, st, d);
      else launch_fwd97<1, false, FWD_STAGES>(grid, block, st, d);
    }
  }
  b2k_count_launch();
}

template <int NC, int STAGES, bool OUT16>
static void launch_inv53(dim3 grid, dim3 block, cudaStream_t st, const DwtLevelDesc* d)
{
  const size_t smem = (size_t)B2K_WARPS_PER_CTA * STAGES * BandStage<NC>::PAIRB + (size_t)B2K_WARPS_PER_CTA * STAGES * sizeof(uint64_t);
  static DeviceOnce once; /* function attributes are per device */
  once.run([&] {
    cudaFuncSetAttribute(k_dwt53_inv<NC, STAGES, OUT16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  k_dwt53_inv<NC, STAGES, OUT16><<<grid, block, smem, st>>>(d);
}

template <int NC, int STAGES>
static void launch_inv97(dim3 grid, dim3 block, cudaStream_t st, const DwtLevelDesc* d)
{
  const size_t smem = (size_t)B2K_WARPS_PER_CTA * STAGES * BandStage<NC>::PAIRB + (size_t)B2K_WARPS_PER_CTA * STAGES * sizeof(uint64_t);
  static DeviceOnce once; /* function attributes are per device */
  once.run([&] {
    cudaFuncSetAttribute(k_dwt97_inv<NC, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  k_dwt97_inv<NC, STAGES><<<grid, block, smem, st>>>(d);
}

/* ---- tiles with NO wavelet level (numres = 1): what is left of the stage is the point transform -- DC shift + RCT / ICT
 * forwards, its inverse + rounding + clamp backwards (mct.cpp L497-636 / L201-391; with one resolution the tile itself is the
 * LL band, TileProcessor.cpp L366-425).  The arithmetic is the level-1 kernels' own (rct_fwd_inplace / ict_fwd_convert,
 * store_rows53 / store_rows97), one sample per thread; the descriptors reuse DwtLevelDesc: in = image samples of the tile
 * component(s), out_c = the tile's place in the coefficient planes. */
template <int NC, bool IRREV, bool FWD>
__global__ void k_point_transform(const DwtLevelDesc* __restrict__ descs, int ndesc)
{
  for(int di = blockIdx.z; di < ndesc; di += gridDim.z)
  {
  const DwtLevelDesc& D = descs[di];
  const int w = D.u1 - D.u0, h = D.v1 - D.v0;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if(x >= w)
    continue;
  for(int y = blockIdx.y; y < h; y += gridDim.y)
  {
  const size_t ii = (size_t)y * D.in_pitch + x, oi = (size_t)y * D.c_pitch + x;
  if(FWD)
  {
    int v[3];
#pragma unroll
    for(int c = 0; c < NC; ++c)
      v[c] = static_cast<const int32_t*>(D.in[c])[ii] + D.shift[c];
    if(!IRREV)
    {
      if(NC == 3)
      {
        const int r = v[0], g = v[NC > 1 ? 1 : 0], b = v[NC > 2 ? 2 : 0];
        v[0] = ((g + g) + b + r) >> 2;
        v[NC > 1 ? 1 : 0] = b - g;
        v[NC > 2 ? 2 : 0] = r - g;
      }
#pragma unroll
      for(int c = 0; c < NC; ++c)
        static_cast<int32_t*>(D.out_c[c])[oi] = v[c];
    }
    else
    {
      float f[3];
      if(NC == 3)
      {
        const float a_r = 0.299f, a_g = 0.587f, a_b = 0.114f;
        const float cb = __fdiv_rn(0.5f, __fsub_rn(1.0f, a_b)), cr = __fdiv_rn(0.5f, __fsub_rn(1.0f, a_r));
        const float r = (float)v[0], g = (float)v[NC > 1 ? 1 : 0], b = (float)v[NC > 2 ? 2 : 0];
        const float yy = __fmaf_rn(a_b, b, __fmaf_rn(a_g, g, __fmul_rn(a_r, r)));
        f[0] = yy;
        f[NC > 1 ? 1 : 0] = __fmul_rn(cb, __fsub_rn(b, yy));
        f[NC > 2 ? 2 : 0] = __fmul_rn(cr, __fsub_rn(r, yy));
      }
      else
        f[0] = (float)v[0];
#pragma unroll
      for(int c = 0; c < NC; ++c)
        static_cast<float*>(D.out_c[c])[oi] = f[c];
    }
  }
  else
  {
    int o[3];
    if(!IRREV)
    {
      int v[3];
#pragma unroll
      for(int c = 0; c < NC; ++c)
        v[c] = static_cast<const int32_t*>(D.out_c[c])[oi];
      if(NC == 3)
      {
        const int yy = v[0], u = v[NC > 1 ? 1 : 0], ww = v[NC > 2 ? 2 : 0];
        const int gg = yy - ((u + ww) >> 2);
        v[0] = ww + gg;
        v[NC > 1 ? 1 : 0] = gg;
        v[NC > 2 ? 2 : 0] = u + gg;
      }
#pragma unroll
      for(int c = 0; c < NC; ++c)
        o[c] = v[c];
    }
    else
    {
      float f[3];
#pragma unroll
      for(int c = 0; c < NC; ++c)
        f[c] = static_cast<const float*>(D.out_c[c])[oi];
      if(NC == 3)
      {
        const float yy = f[0], u = f[NC > 1 ? 1 : 0], ww = f[NC > 2 ? 2 : 0];
        f[0] = __fmaf_rn(ww, 1.402f, yy);
        f[NC > 1 ? 1 : 0] = __fmaf_rn(-ww, 0.71414f, __fmaf_rn(-u, 0.34413f, yy));
        f[NC > 2 ? 2 : 0] = __fmaf_rn(u, 1.772f, yy);
      }
#pragma unroll
      for(int c = 0; c < NC; ++c)
        o[c] = __float2int_rn(f[c]);
    }
#pragma unroll
    for(int c = 0; c < NC; ++c)
      static_cast<int32_t*>(const_cast<void*>(D.in[c]))[ii] = min(max(o[c] - D.shift[c], D.lo[c]), D.hi[c]);
  }
  } /* rows */
  } /* descriptors */
}

void b2k_launch_point_transform(const DwtLevelDesc* d, int ndesc, uint32_t max_w, uint32_t max_h, int nc, bool irreversible,
                                bool forward, cudaStream_t st)
{
  if(ndesc <= 0 || !max_w || !max_h)
    return;
  dim3 grid((max_w + 127) / 128, std::min<uint32_t>(max_h, 65535u), (unsigned)std::min(ndesc, 65535)), block(128);
#define B2K_PT(NC_, IR_, FW_) k_point_transform<NC_, IR_, FW_><<<grid, block, 0, st>>>(d, ndesc)
  if(nc == 3)
  {
    if(irreversible) { if(forward) B2K_PT(3, true, true); else B2K_PT(3, true, false); }
    else { if(forward) B2K_PT(3, false, true); else B2K_PT(3, false, false); }
  }
  else
  {
    if(irreversible) { if(forward) B2K_PT(1, true, true); else B2K_PT(1, true, false); }
    else { if(forward) B2K_PT(1, false, true); else B2K_PT(1, false, false); }
  }
#undef B2K_PT
  b2k_count_launch();
}

void b2k_launch_dwt_inv(const DwtLevelDesc* d, int ndesc, int max_jobs, int nc, bool irreversible, bool out_u16,
                        cudaStream_t st)
{
  if(ndesc <= 0 || max_jobs <= 0)
    return;
  dim3 grid((max_jobs + B2K_WARPS_PER_CTA - 1) / B2K_WARPS_PER_CTA, ndesc), block(B2K_WARPS_PER_CTA * 32);
  if(!irreversible)
  {
    if(out_u16)
    { /* finest level straight into 16-bit containers (widths of 1 are not supported there: the engine checks) */
      if(nc == 3) launch_inv53<3, FWD_STAGES, true>(grid, block, st, d);
      else launch_inv53<1, FWD_STAGES, true>(grid, block, st, d);
    }
    else
    {
      if(nc == 3) launch_inv53<3, FWD_STAGES, false>(grid, block, st, d);
      else launch_inv53<1, FWD_STAGES, false>(grid, block, st, d);
    }
  }
  else
  {
    if(nc == 3) launch_inv97<3, FWD_STAGES>(grid, block, st, d);
    else launch_inv97<1, FWD_STAGES>(grid, block, st, d);
  }
  b2k_count_launch();
}
