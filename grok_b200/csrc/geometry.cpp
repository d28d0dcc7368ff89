/* see geometry.h for the reference citations */
#include "geometry.h"
#include <algorithm>
#include <cmath>

namespace b2k {

TileGrid tile_grid(const b2k_coding& cp)
{
  TileGrid g;
  if(cp.tw == 0 || cp.th == 0)
  { /* untiled: one tile covering the image, anchored at the image origin */
    g.tx0 = cp.x0;
    g.ty0 = cp.y0;
    g.tw = cp.x1 - cp.x0;
    g.th = cp.y1 - cp.y0;
    g.nx = g.ny = 1;
    return g;
  }
  g.tx0 = cp.tx0;
  g.ty0 = cp.ty0;
  g.tw = cp.tw;
  g.th = cp.th;
  g.nx = ceil_div(cp.x1 - cp.tx0, cp.tw);
  g.ny = ceil_div(cp.y1 - cp.ty0, cp.th);
  return g;
}

Rect tile_rect(const b2k_coding& cp, const TileGrid& g, uint32_t t)
{
  const uint32_t p = t % g.nx, q = t / g.nx;
  Rect r;
  r.x0 = std::max<uint64_t>((uint64_t)g.tx0 + (uint64_t)p * g.tw, cp.x0);
  r.y0 = std::max<uint64_t>((uint64_t)g.ty0 + (uint64_t)q * g.th, cp.y0);
  r.x1 = (uint32_t)std::min<uint64_t>((uint64_t)g.tx0 + (uint64_t)(p + 1) * g.tw, cp.x1);
  r.y1 = (uint32_t)std::min<uint64_t>((uint64_t)g.ty0 + (uint64_t)(q + 1) * g.th, cp.y1);
  return r;
}

Rect resolution_rect(const Rect& tc, int numres, int resno)
{
  const uint32_t n = (uint32_t)(numres - 1 - resno);
  return Rect{ceil_div_pow2(tc.x0, n), ceil_div_pow2(tc.y0, n), ceil_div_pow2(tc.x1, n), ceil_div_pow2(tc.y1, n)};
}

static uint32_t band_coord(uint32_t c, uint32_t ndecomp, uint32_t high)
{
  if(ndecomp == 0)
    return c;
  const uint32_t off = (1u << (ndecomp - 1)) * high;
  return c <= off ? 0 : ceil_div_pow2(c - off, ndecomp);
}

Rect band_rect(const Rect& tc, int numres, int resno, int orient)
{
  const uint32_t level = resno == 0 ? (uint32_t)(numres - 1) : (uint32_t)(numres - resno);
  const uint32_t hx = orient & 1, hy = (orient >> 1) & 1;
  return Rect{band_coord(tc.x0, level, hx), band_coord(tc.y0, level, hy), band_coord(tc.x1, level, hx),
              band_coord(tc.y1, level, hy)};
}

/* ---- quantiser ------------------------------------------------------------------------------ */
static const float kBibo53L[34] = {
    1.0000e+00f, 1.5000e+00f, 1.6250e+00f, 1.6875e+00f, 1.6963e+00f, 1.7067e+00f, 1.7116e+00f, 1.7129e+00f,
    1.7141e+00f, 1.7145e+00f, 1.7151e+00f, 1.7152e+00f, 1.7155e+00f, 1.7155e+00f, 1.7156e+00f, 1.7156e+00f,
    1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f,
    1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f, 1.7156e+00f,
    1.7156e+00f, 1.7156e+00f};
static const float kBibo53H[34] = {
    2.0000e+00f, 2.5000e+00f, 2.7500e+00f, 2.8047e+00f, 2.8198e+00f, 2.8410e+00f, 2.8558e+00f, 2.8601e+00f,
    2.8628e+00f, 2.8656e+00f, 2.8662e+00f, 2.8667e+00f, 2.8669e+00f, 2.8670e+00f, 2.8671e+00f, 2.8671e+00f,
    2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f,
    2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f, 2.8671e+00f,
    2.8671e+00f, 2.8671e+00f};
/* sqrt energy gains of the 9/7 synthesis filters, per number of decompositions */
static const float kGain97L[34] = {
    1.0000e+00f, 1.4021e+00f, 2.0304e+00f, 2.9012e+00f, 4.1153e+00f, 5.8245e+00f, 8.2388e+00f, 1.1652e+01f,
    1.6479e+01f, 2.3304e+01f, 3.2957e+01f, 4.6609e+01f, 6.5915e+01f, 9.3217e+01f, 1.3183e+02f, 1.8643e+02f,
    2.6366e+02f, 3.7287e+02f, 5.2732e+02f, 7.4574e+02f, 1.0546e+03f, 1.4915e+03f, 2.1093e+03f, 2.9830e+03f,
    4.2185e+03f, 5.9659e+03f, 8.4371e+03f, 1.1932e+04f, 1.6874e+04f, 2.3864e+04f, 3.3748e+04f, 4.7727e+04f,
    6.7496e+04f, 9.5454e+04f};
static const float kGain97H[34] = {
    1.4425e+00f, 1.9669e+00f, 2.8839e+00f, 4.1475e+00f, 5.8946e+00f, 8.3472e+00f, 1.1809e+01f, 1.6701e+01f,
    2.3620e+01f, 3.3403e+01f, 4.7240e+01f, 6.6807e+01f, 9.4479e+01f, 1.3361e+02f, 1.8896e+02f, 2.6723e+02f,
    3.7792e+02f, 5.3446e+02f, 7.5583e+02f, 1.0689e+03f, 1.5117e+03f, 2.1378e+03f, 3.0233e+03f, 4.2756e+03f,
    6.0467e+03f, 8.5513e+03f, 1.2093e+04f, 1.7103e+04f, 2.4187e+04f, 3.4205e+04f, 4.8373e+04f, 6.8410e+04f,
    9.6747e+04f, 1.3682e+05f};

static void irrev_expn_mant(float delta_b, uint8_t& e, uint16_t& m)
{
  int exp = 0;
  while(delta_b < 1.0f)
  {
    exp++;
    delta_b *= 2.0f;
  }
  int mant = (int)std::round(delta_b * (float)(1 << 11)) - (1 << 11);
  mant = mant < (1 << 11) ? mant : 0x7FF;
  e = (uint8_t)exp;
  m = (uint16_t)mant;
}

std::vector<BandQuant> band_quant(const b2k_coding& cp)
{
  const int D = cp.numres - 1;
  std::vector<BandQuant> q(3 * D + 1);
  std::vector<uint8_t> expn(3 * D + 1);
  std::vector<uint16_t> mant(3 * D + 1, 0);
  int s = 0;
  if(!cp.irreversible)
  {
    const int B = cp.prec + (cp.mct ? 1 : 0);
    float bl = kBibo53L[D];
    int X = (int)std::ceil(std::log(bl * bl * 1.1f) / M_LN2);
    expn[s++] = (uint8_t)(B + X);
    for(int d = D - 1; d >= 0; --d)
    {
      bl = kBibo53L[d + 1];
      const float bh = kBibo53H[d];
      X = (int)std::ceil(std::log(bh * bl * 1.1f) / M_LN2);
      expn[s++] = (uint8_t)(B + X);
      expn[s++] = (uint8_t)(B + X);
      X = (int)std::ceil(std::log(bh * bh * 1.1f) / M_LN2);
      expn[s++] = (uint8_t)(B + X);
    }
  }
  else
  {
    const float base_delta = 1.0f / (float)(1 << (cp.prec + (cp.sgnd ? 1 : 0)));
    const float gl0 = kGain97L[D];
    irrev_expn_mant(base_delta / (gl0 * gl0), expn[s], mant[s]);
    s++;
    for(int d = D; d > 0; --d)
    {
      const float gl = kGain97L[d], gh = kGain97H[d - 1];
      irrev_expn_mant(base_delta / (gl * gh), expn[s], mant[s]);
      expn[s + 1] = expn[s];
      mant[s + 1] = mant[s];
      s += 2;
      irrev_expn_mant(base_delta / (gh * gh), expn[s], mant[s]);
      s++;
    }
  }
  if(cp.qcd_explicit)
    for(int i = 0; i < 3 * D + 1 && i < 97; ++i)
    { /* a foreign stream's QCD (or a caller's own choice) instead of the HT quantiser's tables */
      expn[i] = cp.qcd_expn[i];
      mant[i] = cp.irreversible ? (uint16_t)(cp.qcd_mant[i] & 0x7FF) : 0;
    }
  for(int i = 0; i < 3 * D + 1; ++i)
  {
    const int orient = i == 0 ? 0 : ((i - 1) % 3) + 1;
    const int gain = orient == 0 ? 0 : (orient == 3 ? 2 : 1);
    BandQuant& b = q[i];
    b.expn = expn[i];
    b.mant = mant[i];
    const int k = (int)expn[i] + (int)cp.numgbits - 1;
    b.kmax = (uint8_t)(k > 0 ? k : 0);
    b.step_enc = (float)((1.0 + mant[i] / 2048.0) * std::pow(2.0, (double)((int)cp.prec + gain - (int)expn[i])));
    const int dgain = cp.irreversible ? 0 : gain;
    b.step_dec = (float)((1.0 + mant[i] / 2048.0) * std::pow(2.0, (double)((int)cp.prec + dgain - (int)expn[i])));
  }
  return q;
}

const char* unsupported_reason(const b2k_coding& cp)
{
  if(cp.numcomps < 1 || cp.numcomps > 4)
    return "1..4 components supported";
  if(cp.numres < 1 || cp.numres > 16)
    return "1..16 resolutions supported";
  if(cp.prec < 1 || cp.prec > 16)
    return "precision 1..16 supported";
  if(cp.cblkw_exp < 2 || cp.cblkh_exp < 2 || cp.cblkw_exp > 10 || cp.cblkh_exp > 10 || cp.cblkw_exp + cp.cblkh_exp > 12)
    return "invalid code-block size";
  if(cp.mct && cp.numcomps < 3)
    return "MCT needs three components";
  if(cp.x1 <= cp.x0 || cp.y1 <= cp.y0)
    return "empty image";
  const std::vector<BandQuant> q = band_quant(cp);
  for(const BandQuant& b : q)
    if(b.kmax > 29 || b.kmax < 1)
      return "band bit planes outside the 32-bit HT coder's range";
  return nullptr;
}

void enumerate_tile_blocks(const b2k_coding& cp, uint32_t tile_index, const Rect& tile,
                           const std::vector<BandQuant>& quant, std::vector<b2k_block>& out)
{
  const int numres = cp.numres;
  for(uint16_t comp = 0; comp < cp.numcomps; ++comp)
  {
    const Rect tc = tile; /* dx = dy = 1 */
    for(int resno = 0; resno < numres; ++resno)
    {
      const Rect res = resolution_rect(tc, numres, resno);
      const uint32_t pw = cp.prcw_exp[resno] ? cp.prcw_exp[resno] : 15, ph = cp.prch_exp[resno] ? cp.prch_exp[resno] : 15;
      /* precinct partition of the resolution, then its grid */
      const uint32_t px0 = (res.x0 >> pw) << pw, py0 = (res.y0 >> ph) << ph;
      const uint64_t px1 = (uint64_t)ceil_div_pow2(res.x1, pw) << pw, py1 = (uint64_t)ceil_div_pow2(res.y1, ph) << ph;
      const uint32_t gridw = (uint32_t)((px1 >> pw) - (px0 >> pw)), gridh = (uint32_t)((py1 >> ph) - (py0 >> ph));
      const uint32_t bpw = resno ? pw - 1 : pw, bph = resno ? ph - 1 : ph;
      const uint32_t bpx0 = resno ? px0 >> 1 : px0, bpy0 = resno ? py0 >> 1 : py0;
      const uint32_t cbw = std::min<uint32_t>(cp.cblkw_exp, bpw), cbh = std::min<uint32_t>(cp.cblkh_exp, bph);
      const Rect lower = resno ? resolution_rect(tc, numres, resno - 1) : res;
      const int nbands = resno == 0 ? 1 : 3;
      for(int b = 0; b < nbands; ++b)
      {
        const int orient = resno == 0 ? 0 : b + 1;
        const Rect band = band_rect(tc, numres, resno, orient);
        const BandQuant& bq = quant[band_quant_index(resno, orient)];
        for(uint64_t p = 0; p < (uint64_t)gridw * gridh; ++p)
        {
          Rect prc;
          prc.x0 = bpx0 + (uint32_t)((p % gridw) << bpw);
          prc.y0 = bpy0 + (uint32_t)((p / gridw) << bph);
          prc.x1 = (uint32_t)std::min<uint64_t>((uint64_t)prc.x0 + (1ull << bpw), band.x1);
          prc.y1 = (uint32_t)std::min<uint64_t>((uint64_t)prc.y0 + (1ull << bph), band.y1);
          prc.x0 = std::max(prc.x0, band.x0);
          prc.y0 = std::max(prc.y0, band.y0);
          if(prc.empty())
            continue; /* no code blocks in an empty precinct */
          const uint32_t gx = prc.x0 >> cbw, gy = prc.y0 >> cbh;
          const uint32_t gw = ceil_div_pow2(prc.x1, cbw) - gx, gh = ceil_div_pow2(prc.y1, cbh) - gy;
          for(uint32_t k = 0; k < gw * gh; ++k)
          {
            b2k_block blk{};
            blk.tile = tile_index;
            blk.comp = comp;
            blk.resno = (uint8_t)resno;
            blk.band_index = (uint8_t)b;
            blk.orient = (uint8_t)orient;
            blk.kmax = bq.kmax;
            blk.precno = (uint32_t)p;
            blk.cblkno = k;
            blk.x0 = std::max((gx + k % gw) << cbw, prc.x0);
            blk.y0 = std::max((gy + k / gw) << cbh, prc.y0);
            blk.x1 = (uint32_t)std::min<uint64_t>(((uint64_t)(gx + k % gw) + 1) << cbw, prc.x1);
            blk.y1 = (uint32_t)std::min<uint64_t>(((uint64_t)(gy + k / gw) + 1) << cbh, prc.y1);
            blk.buf_x = blk.x0 - band.x0 + ((resno && (orient & 1)) ? lower.w() : 0);
            blk.buf_y = blk.y0 - band.y0 + ((resno && (orient & 2)) ? lower.h() : 0);
            blk.length = 0;
            blk.offset = 0;
            blk.numbps = 0;
            blk.numpasses = 0;
            blk.stepsize = bq.step_enc;
            out.push_back(blk);
          }
        }
      }
    }
  }
}

} // namespace b2k
