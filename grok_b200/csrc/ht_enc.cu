/*
 * grok_b200/csrc/ht_enc.cu -- HTJ2K (ITU-T T.814) cleanup-pass block ENCODER for sm_100a,
 * one warp per code block, fused with the T1 pre-processing (sign-magnitude conversion and,
 * for the irreversible path, scalar quantisation).
 *
 * Replaces (reference, CPU): T1OJPH::preCompress + compress   t1/part15/CoderOJPH.cpp L121-211
 *                            ojph_encode_codeblock32          t1/part15/coding/ojph_block_encoder.cpp L542-1017
 * Output is byte-identical to that encoder (tests/test_ht_gpu.py, against oracle/ and oracle/_ref).
 *
 * The reference walks quads serially and pushes bits into three byte streams as it goes.
 * Here nothing about a quad depends on coding state, only on neighbouring SAMPLES, so:
 *   lane = quad column (32 quads = 64 sample columns per step), loop over quad rows;
 *   every lane derives rho, exponents, context, kappa, U_q, EMB pattern and looks up its
 *   CxtVLC codeword; warp prefix sums place the variable-length MagSgn / VLC bit strings into
 *   per-warp shared-memory bit rings; rings are drained 32 bytes at a time (one byte per lane)
 *   with the stream's bit-stuffing rule resolved by speculate-and-fix iterations on ballots
 *   (a stuffing event only shifts what follows by one bit, events are rare);
 *   the MEL run-length coder is inherently serial but tiny: the warp builds the event
 *   sequence with ballots/reductions and every lane replays it redundantly (uniform code).
 * The exponent line buffers of the reference (e_val / cx_val, L577-581) become a ping-pong
 * byte line in shared memory; significance is exponent != 0.
 */
#include "b2k_internal.h"
#define HT_TABLE_QUAL static __device__ const
#include "ht_tables.h"

namespace {

constexpr int MS_RING_WORDS = 256;  /* 8192 bits: one 32-quad step adds < 4000, drain leaves < 256 */
constexpr int VLC_RING_WORDS = 64;  /* 2048 bits: one step adds <= 16*30 */
constexpr int MEL_CAP = 256;        /* reference buffer is 192 bytes (L555); more is an error there */

struct WarpShared
{
  uint32_t ms_ring[MS_RING_WORDS];
  uint32_t vlc_ring[VLC_RING_WORDS];
  uint8_t mel[MEL_CAP];
};

__device__ __forceinline__ unsigned lanemask_lt()
{
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

/* append `len` (<= 128) bits held in (lo,hi) at absolute bit position pos of a ring */
__device__ __forceinline__ void ring_put(uint32_t* ring, int ring_words, uint32_t pos, uint64_t lo, uint64_t hi, int len)
{
  if(len <= 0)
    return;
  const int sh = pos & 31;
  uint32_t w = pos >> 5;
  /* 160-bit shifted value in five 32-bit words */
  uint32_t v[5];
  const uint32_t a0 = (uint32_t)lo, a1 = (uint32_t)(lo >> 32), a2 = (uint32_t)hi, a3 = (uint32_t)(hi >> 32);
  v[0] = a0 << sh;
  v[1] = __funnelshift_l(a0, a1, sh);
  v[2] = __funnelshift_l(a1, a2, sh);
  v[3] = __funnelshift_l(a2, a3, sh);
  v[4] = sh ? (a3 >> (32 - sh)) : 0u;
  const int nwords = (sh + len + 31) >> 5;
#pragma unroll
  for(int i = 0; i < 5; ++i)
    if(i < nwords && v[i])
      atomicOr(&ring[(w + i) & (ring_words - 1)], v[i]);
}

/* 15 bits starting at absolute bit position pos; bits at or beyond `tail` read as `fill`
   (MASK = false: the caller guarantees pos + 15 <= tail) */
template <bool MASK = true>
__device__ __forceinline__ uint32_t ring_get15(const uint32_t* ring, int ring_words, uint32_t pos, uint32_t tail,
                                               uint32_t fill)
{
  const uint32_t w = pos >> 5;
  const int sh = pos & 31;
  const uint32_t lo = ring[w & (ring_words - 1)], hi = ring[(w + 1) & (ring_words - 1)];
  uint32_t v = __funnelshift_r(lo, hi, sh) & 0x7FFFu;
  const int avail = (int)(tail - pos);
  if(MASK && avail < 15)
  {
    const uint32_t keep = avail <= 0 ? 0u : ((1u << avail) - 1u);
    v = (v & keep) | ((fill ? 0x7FFFu : 0u) & ~keep);
  }
  return v;
}

/* zero ring words fully below new_head */
__device__ __forceinline__ void ring_release(uint32_t* ring, int ring_words, uint32_t old_head, uint32_t new_head, int lane)
{
  const uint32_t w0 = old_head >> 5, w1 = new_head >> 5;
  if(w1 == w0)
    return;
  for(uint32_t w = w0 + lane; w < w1; w += 32)
    ring[w & (ring_words - 1)] = 0;
}

/* ---------------------------------------------------------------------------------------------
 * Drain up to 32 bytes of the forward MagSgn stream (ms_encode, ojph_block_encoder.cpp L470-491):
 * a byte following 0xFF carries 7 bits.  Returns number of bytes written.
 * final: pad the tail with 1s and emit the last partial byte too (ms_terminate L516-535; the
 * caller drops a trailing 0xFF).
 * ------------------------------------------------------------------------------------------- */
template <bool FINAL>
__device__ __forceinline__ int ms_drain32(uint32_t* ring, uint32_t& head, uint32_t tail, bool& last_ff, uint8_t* out,
                                          int lane, uint32_t& last_byte)
{
  const bool final = FINAL;
  unsigned ffmask = 0;
  uint32_t byte = 0, start = 0;
  int nbits = 8;
  for(int it = 0; it < 34; ++it)
  {
    /* lanes below i that produced 0xFF each shift everything after them by one bit */
    const unsigned prevff = (ffmask << 1) | (last_ff ? 1u : 0u); /* bit i: byte i-1 is 0xFF */
    nbits = ((prevff >> lane) & 1u) ? 7 : 8;
    const int stuffed_before = __popc(prevff & lanemask_lt()) ; /* 7-bit bytes among lanes < i */
    start = head + 8u * lane - (uint32_t)stuffed_before;
    /* non-final drains run with >= 256 bits queued, but lane 31's window can still poke past the tail */
    const uint32_t raw = ring_get15<true>(ring, MS_RING_WORDS, start, tail, 1u);
    byte = raw & (nbits == 7 ? 0x7Fu : 0xFFu);
    const unsigned nf = __ballot_sync(0xffffffffu, byte == 0xFFu);
    if(nf == ffmask)
      break;
    ffmask = nf;
  }
  /* complete bytes: all bits available, or (final) at least one real bit */
  const bool complete = final ? (start < tail) : (start + nbits <= tail);
  const unsigned cm = __ballot_sync(0xffffffffu, complete);
  const int nb = (cm == 0xffffffffu) ? 32 : (__ffs(~cm) - 1);
  if(lane < nb)
    out[lane] = (uint8_t)byte;
  /* advance */
  const uint32_t endpos = start + nbits;
  const uint32_t new_head = nb ? __shfl_sync(0xffffffffu, endpos, nb - 1) : head;
  if(nb)
  {
    last_byte = __shfl_sync(0xffffffffu, byte, nb - 1);
    last_ff = (last_byte == 0xFFu);
  }
  head = new_head; /* consumed ring words are zeroed by the caller, once per step (ring_release) */
  return nb;
}

/* ---------------------------------------------------------------------------------------------
 * Drain exactly 128 bytes of the forward MagSgn stream, 4 bytes per lane (requires >= 1024 raw bits
 * queued).  A lane walks its own four bytes serially (a byte after 0xFF carries 7 bits); how many
 * stuffed bytes sit in the lanes below -- which shifts the lane's window one bit each -- is
 * resolved with the same speculate-and-fix iteration on ballots as the 32-byte drain.
 * ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void ms_drain128(const uint32_t* ring, uint32_t& head, bool& last_ff, uint8_t* out, int lane)
{
  unsigned m1 = 0, m2 = 0, lf = 0; /* lanes with >=1 / >=2 seven-bit bytes; lanes whose 4th byte is 0xFF */
  uint32_t word = 0;
  for(int it = 0; it < 34; ++it)
  {
    const unsigned below = lanemask_lt();
    const uint32_t start = head + 32u * lane - (uint32_t)(__popc(m1 & below) + __popc(m2 & below));
    const bool pff = lane == 0 ? last_ff : (((lf >> (lane - 1)) & 1u) != 0);
    const uint32_t wi = start >> 5;
    const int sh = start & 31;
    uint32_t raw = __funnelshift_r(ring[wi & (MS_RING_WORDS - 1)], ring[(wi + 1) & (MS_RING_WORDS - 1)], sh);
    int sev = 0;
    bool f = pff;
    word = 0;
#pragma unroll
    for(int j = 0; j < 4; ++j)
    {
      const uint32_t b = raw & (f ? 0x7Fu : 0xFFu);
      raw >>= f ? 7 : 8;
      sev += f ? 1 : 0;
      f = (b == 0xFFu);
      word |= b << (8 * j);
    }
    const unsigned n1 = __ballot_sync(0xffffffffu, sev >= 1), n2 = __ballot_sync(0xffffffffu, sev >= 2),
                   nf = __ballot_sync(0xffffffffu, f);
    if(n1 == m1 && n2 == m2 && nf == lf)
      break;
    m1 = n1; m2 = n2; lf = nf;
  }
  *reinterpret_cast<uint32_t*>(out + 4 * lane) = word;
  head += 1024u - (uint32_t)(__popc(m1) + __popc(m2));
  last_ff = (lf >> 31) & 1u;
}

/* ---------------------------------------------------------------------------------------------
 * Drain up to 32 bytes of the backward VLC stream (vlc_encode L378-410): a byte that follows one
 * > 0x8F and whose first 7 bits are all ones is emitted as 0x7F and carries 7 bits.  Only
 * complete bytes are written; byte k of the stream goes to out_last[-k].
 * ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int vlc_drain32(uint32_t* ring, uint32_t& head, uint32_t tail, uint32_t& prev_byte,
                                           uint8_t* out_first, int lane)
{
  unsigned smask = 0; /* bit i: byte i is a 7-bit (stuffed) byte */
  uint32_t byte = 0, start = 0;
  int nbits = 8;
  for(int it = 0; it < 34; ++it)
  {
    const int before = __popc(smask & lanemask_lt());
    start = head + 8u * lane - (uint32_t)before;
    const uint32_t raw = ring_get15(ring, VLC_RING_WORDS, start, tail, 0u);
    const uint32_t b8 = raw & 0xFFu;
    /* tentative value of the previous byte */
    uint32_t pb = __shfl_up_sync(0xffffffffu, byte, 1);
    if(lane == 0)
      pb = prev_byte;
    const bool stuffed = (pb > 0x8Fu) && ((raw & 0x7Fu) == 0x7Fu);
    nbits = stuffed ? 7 : 8;
    const uint32_t nbyte = stuffed ? 0x7Fu : b8;
    const unsigned ns = __ballot_sync(0xffffffffu, stuffed);
    const bool same = __all_sync(0xffffffffu, nbyte == byte) && ns == smask;
    byte = nbyte;
    smask = ns;
    if(same && it > 0)
      break;
  }
  const bool complete = start + nbits <= tail;
  const unsigned cm = __ballot_sync(0xffffffffu, complete);
  const int nb = (cm == 0xffffffffu) ? 32 : (__ffs(~cm) - 1);
  if(lane < nb)
    *(out_first - lane) = (uint8_t)byte;
  const uint32_t endpos = start + nbits;
  const uint32_t new_head = nb ? __shfl_sync(0xffffffffu, endpos, nb - 1) : head;
  if(nb)
    prev_byte = __shfl_sync(0xffffffffu, byte, nb - 1);
  head = new_head;
  return nb;
}

/* ---- MEL coder state (mel_struct, L273-345), identical in every lane ------------------------ */
struct Mel
{
  int rem, tmp, run, k, thr, pos;
};
__device__ __forceinline__ int mel_exp(int k) { return (int)((0x58da489200ull >> (3 * k)) & 7ull); }
/* exponents {0,0,0,1,1,1,2,2,2,3,3,4,5}: 3 bits each, k=0 lowest */

__device__ __forceinline__ void mel_emit(Mel& m, int v, uint8_t* buf, int lane)
{
  m.tmp = (m.tmp << 1) + v;
  if(--m.rem == 0)
  {
    if(lane == 0 && m.pos < MEL_CAP)
      buf[m.pos] = (uint8_t)m.tmp;
    m.pos++;
    m.rem = (m.tmp == 0xFF) ? 7 : 8;
    m.tmp = 0;
  }
}
__device__ __forceinline__ void mel_zeros(Mel& m, int n, uint8_t* buf, int lane)
{
  while(n > 0)
  {
    const int need = m.thr - m.run;
    if(n >= need)
    {
      mel_emit(m, 1, buf, lane);
      m.run = 0;
      m.k = min(12, m.k + 1);
      m.thr = 1 << mel_exp(m.k);
      n -= need;
    }
    else
    {
      m.run += n;
      n = 0;
    }
  }
}
__device__ __forceinline__ void mel_one(Mel& m, uint8_t* buf, int lane)
{
  mel_emit(m, 0, buf, lane);
  for(int t = mel_exp(m.k); t > 0;)
    mel_emit(m, (m.run >> --t) & 1, buf, lane);
  m.run = 0;
  m.k = max(0, m.k - 1);
  m.thr = 1 << mel_exp(m.k);
}

/* UVLC codeword (uvlc_tbl, L196-256): prefix[2:0] | prefix_len<<3 | suffix<<6 | suffix_len<<11 for
   u = 0..32 (u==0: nothing; 1: "1"; 2: "01"; 3,4: "001"+1 bit; 5..32: "000"+5 bits) */
__constant__ uint16_t UVLC_LUT[33] = {0x0000, 0x0009, 0x0012, 0x081C, 0x085C, 0x2818, 0x2858, 0x2898, 0x28D8, 0x2918, 0x2958, 0x2998, 0x29D8, 0x2A18, 0x2A58, 0x2A98, 0x2AD8, 0x2B18, 0x2B58, 0x2B98, 0x2BD8, 0x2C18, 0x2C58, 0x2C98, 0x2CD8, 0x2D18, 0x2D58, 0x2D98, 0x2DD8, 0x2E18, 0x2E58, 0x2E98, 0x2ED8};
__device__ __forceinline__ void uvlc_bits(int u, uint32_t& pre, int& prel, uint32_t& suf, int& sufl)
{
  const uint32_t t = UVLC_LUT[u > 32 ? 32 : u];
  pre = t & 7u;
  prel = (int)((t >> 3) & 7u);
  suf = (t >> 6) & 31u;
  sufl = (int)(t >> 11);
}

template <typename T>
__device__ __forceinline__ T warp_excl_scan(T v, int lane, T& total)
{
  T x = v;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const T y = __shfl_up_sync(0xffffffffu, x, o);
    if(lane >= o)
      x += y;
  }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

template <bool IRREV>
__global__ void __launch_bounds__(B2K_WARPS_PER_CTA * 32)
    k_ht_encode(const HtBlockDesc* __restrict__ blocks, HtBlockOut* __restrict__ outs, uint8_t* __restrict__ scratch,
                uint32_t nblocks, uint32_t line_entries)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint16_t* tbl0 = reinterpret_cast<uint16_t*>(smem_raw);
  uint16_t* tbl1 = tbl0 + 2048;
  WarpShared* wsh_all = reinterpret_cast<WarpShared*>(smem_raw + 2 * 2048 * sizeof(uint16_t));
  uint16_t* lines_all = reinterpret_cast<uint16_t*>(wsh_all + B2K_WARPS_PER_CTA);

  for(int i = threadIdx.x; i < 2048; i += blockDim.x)
  {
    tbl0[i] = HT_ENC_VLC0[i];
    tbl1[i] = HT_ENC_VLC1[i];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bidx = blockIdx.x * B2K_WARPS_PER_CTA + warp;
  if(bidx >= nblocks)
    return;
  const HtBlockDesc B = blocks[bidx];
  WarpShared& S = wsh_all[warp];
  /* exponent lines: entry q+1 holds the bottom-row exponents of columns 2q (low byte), 2q+1 */
  uint16_t* line[2] = {lines_all + (size_t)warp * 2 * line_entries, lines_all + (size_t)warp * 2 * line_entries + line_entries};

  for(int i = lane; i < MS_RING_WORDS; i += 32)
    S.ms_ring[i] = 0;
  for(int i = lane; i < VLC_RING_WORDS; i += 32)
    S.vlc_ring[i] = 0;
  for(uint32_t i = lane; i < 2 * line_entries; i += 32)
    line[0][i] = 0;
  __syncwarp();

  const int w = B.w, h = B.h;
  const int nq = (w + 1) >> 1;
  const int kmax = B.kmax;
  const int shift = 30 - kmax; /* CoderOJPH.cpp L131: 31 - (k_msbs + 1) */
  uint8_t* slot = scratch + B.slot_off;
  uint8_t* slot_last = slot + B.slot_cap - 1; /* VLC byte k lives at slot_last[-k] (vlc_init L364-375) */

  /* stream states */
  uint32_t ms_head = 0, ms_tail = 0, ms_out = 0, ms_lastbyte = 0;
  bool ms_lastff = false;
  uint32_t vlc_head = 0, vlc_tail = 4, vlc_out = 1, vlc_prev = 0xFF; /* 4 one-bits, virtual previous byte > 0x8F */
  if(lane == 0)
  {
    S.vlc_ring[0] = 0xF;
    *slot_last = 0xFF;
  }
  Mel mel = {8, 0, 0, 0, 1, 0};
  __syncwarp();

  const float fscale = (float)(1u << shift);

  /* flattened (quad row, 32-quad chunk) steps so the NEXT step's four samples are already in flight
     while this step is coded */
  const int nch = (nq + 31) >> 5, nsteps = ((h + 1) >> 1) * nch;
  auto fetch = [&](int step, uint32_t (&raw)[4]) {
    const int yy0 = 2 * (step / nch), xq = 2 * ((step % nch) * 32 + lane);
#pragma unroll
    for(int i = 0; i < 4; ++i)
    {
      const int xx = xq + (i >> 1), yy = yy0 + (i & 1);
      raw[i] = 0;
      if(step < nsteps && xx < w && yy < h)
        raw[i] = __ldg(reinterpret_cast<const uint32_t*>(B.coef) + ((size_t)yy * B.pitch + xx));
    }
  };
  uint32_t cur[4], nxt[4];
  fetch(0, cur);
  int rho_carry = 0;
  for(int step = 0; step < nsteps; ++step)
  {
    const int y = 2 * (step / nch), q0 = (step % nch) * 32;
    const uint16_t* labove = line[(y >> 1) & 1];
    uint16_t* lcur = line[((y >> 1) & 1) ^ 1];
    if(q0 == 0)
      rho_carry = 0;
    fetch(step + 1, nxt);
    {
      const int q = q0 + lane, x = 2 * q;
      const bool qv = q < nq;
      /* ---- samples -> (mu, sign) ---- */
      int rho = 0, emax = 0;
      int e[4];
      uint32_t sv[4];
#pragma unroll
      for(int i = 0; i < 4; ++i)
      {
        e[i] = 0;
        sv[i] = 0;
        uint32_t mu, sgn;
        if(!IRREV)
        {
          const int32_t v = (int32_t)cur[i];
          sgn = (uint32_t)v >> 31;
          mu = (uint32_t)(v < 0 ? -v : v);
          /* bits of |v| above Kmax are shifted out by the reference's `mag << shift; t + t` */
          mu &= (1u << (kmax + 1)) - 1u;
        }
        else
        { /* CoderOJPH.cpp L166-180 */
          const float f = __uint_as_float(cur[i]);
          const int32_t t = __float2int_rz(__fmul_rn(__fmul_rn(f, B.quant), fscale));
          sgn = (uint32_t)t >> 31;
          const uint32_t m = (uint32_t)(t < 0 ? -t : t);
          mu = ((m + m) >> shift) >> 1;
        }
        if(mu)
        {
          rho |= 1 << i;
          e[i] = 32 - __clz(2 * mu - 1);
          emax = max(emax, e[i]);
          sv[i] = 2 * (mu - 1) + sgn;
        }
        cur[i] = nxt[i];
      }
      /* ---- neighbours ---- */
      int rho_left = __shfl_up_sync(0xffffffffu, rho, 1);
      if(lane == 0)
        rho_left = rho_carry;
      rho_carry = __shfl_sync(0xffffffffu, rho, 31);
      int cq, kappa = 1;
      if(y == 0)
        cq = (rho_left >> 1) | (rho_left & 1);
      else
      {
        /* exponents of row y-1 at columns x-1, x, x+1, x+2 */
        const uint32_t a = qv ? labove[q] : 0u, b = qv ? labove[q + 1] : 0u, c = qv ? labove[q + 2] : 0u;
        const int em1 = (int)(a >> 8), e0 = (int)(b & 0xFF), e1 = (int)(b >> 8), e2 = (int)(c & 0xFF);
        cq = ((em1 | e0) ? 1 : 0) | ((rho_left & 0xC) ? 2 : 0) | ((e1 | e2) ? 4 : 0);
        const int max_e = max(max(em1, e0), max(e1, e2)) - 1;
        kappa = (rho & (rho - 1)) ? max(1, max_e) : 1;
      }
      if(qv)
        lcur[q + 1] = (uint16_t)(e[1] | (e[3] << 8));
      const int Uq = max(emax, kappa);
      const int uq = Uq - kappa;
      int eps = 0;
      if(uq > 0)
      {
#pragma unroll
        for(int i = 0; i < 4; ++i)
          eps |= (e[i] == emax) << i;
      }
      const uint32_t tuple = qv ? (uint32_t)((y ? tbl1 : tbl0)[(cq << 8) + (rho << 4) + eps]) : 0u;
      const int cwd_len = (tuple >> 4) & 7;
      const uint32_t cwd = tuple >> 8;

      /* ---- MagSgn bits of this quad ---- */
      uint64_t mlo = 0, mhi = 0;
      int mlen = 0;
#pragma unroll
      for(int i = 0; i < 4; ++i)
      {
        const int m = (rho & (1 << i)) ? Uq - (int)((tuple >> i) & 1) : 0;
        if(m > 0)
        {
          const uint64_t bits = (uint64_t)(sv[i] & (m >= 32 ? 0xFFFFFFFFu : ((1u << m) - 1u)));
          if(mlen < 64)
          {
            mlo |= bits << mlen;
            if(mlen + m > 64)
              mhi |= bits >> (64 - mlen);
          }
          else
            mhi |= bits << (mlen - 64);
          mlen += m;
        }
      }

      /* ---- VLC bits of the quad pair (even lane builds them) ---- */
      const uint32_t cwd1 = __shfl_down_sync(0xffffffffu, cwd, 1);
      const int len1 = __shfl_down_sync(0xffffffffu, cwd_len, 1);
      const int u1raw = __shfl_down_sync(0xffffffffu, uq, 1);
      const bool pair_has1 = (q + 1) < nq;
      uint32_t vbits = 0;
      int vlen = 0;
      int u0 = uq, u1 = pair_has1 ? u1raw : 0;
      bool pair_ev = false, pair_ev_val = false;
      if(qv && !(lane & 1))
      {
        vbits = cwd;
        vlen = cwd_len;
        if(pair_has1)
        {
          vbits |= cwd1 << vlen;
          vlen += len1;
        }
        uint32_t p0, s0, p1, s1;
        int pl0, sl0, pl1, sl1;
        if(y == 0)
        { /* L750-785 */
          if(u0 > 0 && u1 > 0)
          {
            pair_ev = true;
            pair_ev_val = min(u0, u1) > 2;
          }
          if(u0 > 2 && u1 > 2)
          {
            uvlc_bits(u0 - 2, p0, pl0, s0, sl0);
            uvlc_bits(u1 - 2, p1, pl1, s1, sl1);
          }
          else if(u0 > 2 && u1 > 0)
          {
            uvlc_bits(u0, p0, pl0, s0, sl0);
            p1 = (uint32_t)(u1 - 1); pl1 = 1; s1 = 0; sl1 = 0;
          }
          else
          {
            uvlc_bits(u0, p0, pl0, s0, sl0);
            uvlc_bits(u1, p1, pl1, s1, sl1);
          }
        }
        else
        {
          uvlc_bits(u0, p0, pl0, s0, sl0);
          uvlc_bits(u1, p1, pl1, s1, sl1);
        }
        vbits |= p0 << vlen; vlen += pl0;
        vbits |= p1 << vlen; vlen += pl1;
        vbits |= s0 << vlen; vlen += sl0;
        vbits |= s1 << vlen; vlen += sl1;
      }
      /* one warp scan places both bit strings: MagSgn length (< 2^16 per step) in the low half,
         VLC length (<= 16*30) in the high half */
      uint32_t both_total;
      const uint32_t both_off = warp_excl_scan<uint32_t>((uint32_t)mlen | ((uint32_t)vlen << 16), lane, both_total);
      ring_put(S.ms_ring, MS_RING_WORDS, ms_tail + (both_off & 0xFFFFu), mlo, mhi, mlen);
      ms_tail += both_total & 0xFFFFu;
      ring_put(S.vlc_ring, VLC_RING_WORDS, vlc_tail + (both_off >> 16), (uint64_t)vbits, 0ull, vlen);
      vlc_tail += both_total >> 16;

      /* ---- MEL events, in coding order: quad 2p, quad 2p+1, pair p (L664-665, L750-751) ---- */
      {
        const bool ev = qv && cq == 0;
        const int pos = 3 * (lane >> 1) + (lane & 1);
        uint32_t hlo = 0, hhi = 0, vlo = 0, vhi = 0;
        if(ev)
        {
          if(pos < 32) hlo |= 1u << pos; else hhi |= 1u << (pos - 32);
          if(rho) { if(pos < 32) vlo |= 1u << pos; else vhi |= 1u << (pos - 32); }
        }
        if(pair_ev)
        {
          const int pp = 3 * (lane >> 1) + 2;
          if(pp < 32) hlo |= 1u << pp; else hhi |= 1u << (pp - 32);
          if(pair_ev_val) { if(pp < 32) vlo |= 1u << pp; else vhi |= 1u << (pp - 32); }
        }
        hlo = __reduce_or_sync(0xffffffffu, hlo);
        hhi = __reduce_or_sync(0xffffffffu, hhi);
        vlo = __reduce_or_sync(0xffffffffu, vlo);
        vhi = __reduce_or_sync(0xffffffffu, vhi);
        uint64_t has = ((uint64_t)hhi << 32) | hlo;
        const uint64_t val = ((uint64_t)vhi << 32) | vlo;
        while(has)
        {
          const uint64_t ones = has & val;
          if(ones == 0)
          {
            mel_zeros(mel, __popcll(has), S.mel, lane);
            break;
          }
          const int b = __ffsll((long long)ones) - 1;
          mel_zeros(mel, __popcll(has & ((1ull << b) - 1ull)), S.mel, lane);
          mel_one(mel, S.mel, lane);
          has &= ~((2ull << b) - 1ull);
        }
      }
      __syncwarp();

      /* ---- drain full 32-byte windows, then zero the ring words that were consumed ---- */
      {
        const uint32_t h0 = ms_head, v0 = vlc_head;
        while(ms_tail - ms_head >= 1024u)
        {
          ms_drain128(S.ms_ring, ms_head, ms_lastff, slot + ms_out, lane);
          ms_out += 128u;
        }
        while(vlc_tail - vlc_head >= 256u)
          vlc_out += (uint32_t)vlc_drain32(S.vlc_ring, vlc_head, vlc_tail, vlc_prev, slot_last - vlc_out, lane);
        __syncwarp();
        ring_release(S.ms_ring, MS_RING_WORDS, h0, ms_head, lane);
        ring_release(S.vlc_ring, VLC_RING_WORDS, v0, vlc_head, lane);
        __syncwarp();
      }
    }
  }

  /* ---- terminate MagSgn (ms_terminate L516-535) ---- */
  while(ms_head < ms_tail)
    ms_out += (uint32_t)ms_drain32<true>(S.ms_ring, ms_head, ms_tail, ms_lastff, slot + ms_out, lane, ms_lastbyte);
  if(ms_out > 0 && ms_lastff)
    ms_out--; /* a final 0xFF is not written (padded partial byte) or is taken back (L533-534) */

  /* ---- VLC: flush complete bytes, keep the partial one for the MEL/VLC fusion ---- */
  for(;;)
  {
    const int n = vlc_drain32(S.vlc_ring, vlc_head, vlc_tail, vlc_prev, slot_last - vlc_out, lane);
    vlc_out += (uint32_t)n;
    if(n < 32)
      break;
  }
  const int vused = (int)(vlc_tail - vlc_head);
  const int vtmp = (int)ring_get15(S.vlc_ring, VLC_RING_WORDS, vlc_head, vlc_tail, 0u) & 0xFF;

  /* ---- terminate_mel_vlc (L412-444) ---- */
  if(mel.run > 0)
    mel_emit(mel, 1, S.mel, lane);
  {
    const int mtmp = (mel.tmp << mel.rem) & 0xFFFF;
    const int mel_mask = (0xFF << mel.rem) & 0xFF;
    const int vlc_mask = vused ? (0xFF >> (8 - vused)) : 0;
    if((mel_mask | vlc_mask) != 0)
    {
      const int fuse = mtmp | vtmp;
      if(((((fuse ^ mtmp) & mel_mask) | ((fuse ^ vtmp) & vlc_mask)) == 0) && fuse != 0xFF && vlc_out > 1)
      {
        if(lane == 0 && mel.pos < MEL_CAP)
          S.mel[mel.pos] = (uint8_t)fuse;
        mel.pos++;
      }
      else
      {
        if(lane == 0 && mel.pos < MEL_CAP)
          S.mel[mel.pos] = (uint8_t)mtmp;
        mel.pos++;
        if(lane == 0)
          *(slot_last - vlc_out) = (uint8_t)vtmp;
        vlc_out++;
      }
    }
  }
  __syncwarp();
  /* MEL bytes follow the MagSgn bytes */
  for(int i = lane; i < mel.pos && i < MEL_CAP; i += 32)
    slot[ms_out + i] = S.mel[i];
  /* interface locator word (L1009-1014) */
  const uint32_t scup = (uint32_t)mel.pos + vlc_out;
  __syncwarp();
  if(lane == 0)
  {
    slot_last[0] = (uint8_t)(scup >> 4);
    slot_last[-1] = (uint8_t)((slot_last[-1] & 0xF0) | (scup & 0xF));
    HtBlockOut o;
    o.ms_len = ms_out;
    o.mel_len = (uint32_t)mel.pos;
    o.vlc_len = vlc_out;
    o.total = ms_out + (uint32_t)mel.pos + vlc_out;
    if(mel.pos > 192 || o.total > B.slot_cap)
      o.total = 0xFFFFFFFFu; /* the reference raises "mel encoder's buffer is full" here */
    outs[bidx] = o;
  }
}

/* lengths -> exclusive byte offsets (single CTA scan; nblocks is ~5e4) */
__global__ void k_scan_lengths(const HtBlockOut* __restrict__ outs, uint64_t* __restrict__ offsets, uint32_t n)
{
  __shared__ uint64_t partial[1024];
  __shared__ uint64_t carry;
  if(threadIdx.x == 0)
    carry = offsets[0]; /* running base: 0 for the first range, the previous range's end otherwise */
  __syncthreads();
  for(uint32_t base = 0; base < n; base += 1024)
  {
    const uint32_t i = base + threadIdx.x;
    uint64_t v = 0;
    if(i < n)
    {
      const uint32_t t = outs[i].total;
      v = (t == 0xFFFFFFFFu) ? 0 : t;
    }
    partial[threadIdx.x] = v;
    __syncthreads();
    for(int o = 1; o < 1024; o <<= 1)
    {
      uint64_t t = threadIdx.x >= o ? partial[threadIdx.x - o] : 0;
      __syncthreads();
      partial[threadIdx.x] += t;
      __syncthreads();
    }
    if(i < n)
      offsets[i] = carry + partial[threadIdx.x] - v;
    __syncthreads();
    if(threadIdx.x == 1023)
      carry += partial[1023];
    __syncthreads();
  }
  if(threadIdx.x == 0)
    offsets[n] = carry;
}

/* compaction: MagSgn|MEL from the slot head, VLC from the slot tail (warp per block) */
__global__ void k_ht_gather(const HtBlockDesc* __restrict__ blocks, const HtBlockOut* __restrict__ outs,
                            const uint64_t* __restrict__ offsets, const uint8_t* __restrict__ scratch,
                            uint8_t* __restrict__ bytes, uint32_t nblocks, uint64_t cap)
{
  const uint32_t bidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if(bidx >= nblocks)
    return;
  const HtBlockOut o = outs[bidx];
  if(o.total == 0xFFFFFFFFu || offsets[bidx] + o.total > cap)
    return; /* arena too small: the host notices from the offsets and re-gathers */
  const uint8_t* slot = scratch + blocks[bidx].slot_off;
  uint8_t* dst = bytes + offsets[bidx];
  const uint32_t front = o.ms_len + o.mel_len;
  for(uint32_t i = lane; i < front; i += 32)
    dst[i] = slot[i];
  const uint8_t* vsrc = slot + blocks[bidx].slot_cap - o.vlc_len;
  for(uint32_t i = lane; i < o.vlc_len; i += 32)
    dst[front + i] = vsrc[i];
}

} /* namespace */

void b2k_launch_ht_encode(const HtBlockDesc* d_blocks, HtBlockOut* d_out, uint8_t* d_scratch, uint32_t nblocks,
                          uint32_t max_w, bool irreversible, cudaStream_t st)
{
  if(!nblocks)
    return;
  const uint32_t line_entries = ((max_w + 1) / 2 + 4 + 1) & ~1u;
  const size_t smem = 2 * 2048 * sizeof(uint16_t) + B2K_WARPS_PER_CTA * sizeof(WarpShared) +
                      (size_t)B2K_WARPS_PER_CTA * 2 * line_entries * sizeof(uint16_t);
  static DeviceOnce once; /* function attributes are per device */
  once.run([&] {
    cudaFuncSetAttribute(k_ht_encode<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_ht_encode<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  });
  const uint32_t grid = (nblocks + B2K_WARPS_PER_CTA - 1) / B2K_WARPS_PER_CTA;
  if(irreversible)
    k_ht_encode<true><<<grid, B2K_WARPS_PER_CTA * 32, smem, st>>>(d_blocks, d_out, d_scratch, nblocks, line_entries);
  else
    k_ht_encode<false><<<grid, B2K_WARPS_PER_CTA * 32, smem, st>>>(d_blocks, d_out, d_scratch, nblocks, line_entries);
  b2k_count_launch();
}

void b2k_launch_scan_lengths(const HtBlockOut* d_out, uint64_t* d_offsets, uint32_t nblocks, cudaStream_t st)
{
  k_scan_lengths<<<1, 1024, 0, st>>>(d_out, d_offsets, nblocks);
  b2k_count_launch();
}

void b2k_launch_ht_gather(const HtBlockDesc* d_blocks, const HtBlockOut* d_out, const uint64_t* d_offsets,
                          const uint8_t* d_scratch, uint8_t* d_bytes, uint32_t nblocks, uint64_t cap, cudaStream_t st)
{
  if(!nblocks)
    return;
  const uint32_t threads = 256, wpb = threads / 32;
  k_ht_gather<<<(nblocks + wpb - 1) / wpb, threads, 0, st>>>(d_blocks, d_out, d_offsets, d_scratch, d_bytes, nblocks, cap);
  b2k_count_launch();
}
