/*
 * grok_b200/csrc/plugin_decode.cpp -- the stock decode entry point, plugin_decompress().
 *
 * Protocol (host side: CodeStreamDecompress.cpp L199-271, grok.cpp L1882-2015): the PLUGIN drives and
 * calls the host back four times with decompress_flags
 *   HEADER  -> host parses the main header, fills header_info + image, and calls OUR
 *              init_decompressors_func(header_info, image): we size the engine and allocate the tile tree
 *              (every code block gets a buffer the host will copy its bytes into);
 *   T2      -> host parses packets only and fills, per block, compressedData / compressedDataLength /
 *              numBitPlanes / numPasses (decompress_synch_plugin_with_host, TileProcessor.cpp L157-227);
 *   POST_T1 -> we have decoded: host copies the int32 planes out (pluginStoreDecodedImage L243-271);
 *   CLEAN   -> both sides drop per-tile state.
 * Return: 0 decoded, 1 not handled (host decodes on the CPU), <0 error (plugin_accelerate.h L32-36).
 * Eligibility mirrors the host's own (single tile, no reduce/window/layers: L199-205) plus: HT cleanup
 * pass only, one segment per block, this engine's quantiser (Grok's HT defaults).
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "plugin_decode_abi.h"
#include "geometry.h"

using namespace b2k;

extern b2k_engine* b2k_plugin_engine(void); /* plugin.cpp: the engine plugin_init created (or creates it) */
extern void b2k_plugin_free_tree(gpup_tile* T);

namespace {

struct DecodeCtx
{
  b2k_coding cp{};
  bool ready = false;
  int32_t rc = 0;
  std::vector<b2k_block> blocks;
  std::vector<gpup_code_block*> cb_of; /* tree node of every enumerated block */
  std::vector<gpup_band*> bands;       /* comp-major, then resolution, then band */
  gpup_tile* tree = nullptr;
  uint8_t* slab = nullptr; /* host-visible block buffers the T2 callback fills */
  uint64_t slab_bytes = 0;
};
thread_local DecodeCtx* g_ctx = nullptr;

int floor_log2_u32(uint32_t v)
{
  int l = 0;
  while(v >>= 1)
    ++l;
  return l;
}

uint32_t block_capacity(uint32_t w, uint32_t h, uint32_t kmax)
{
  const uint64_t samples = (uint64_t)w * h, quads = (uint64_t)((w + 1) / 2) * ((h + 1) / 2);
  return (uint32_t)(((samples * (kmax + 2) + 6) / 7 + (quads * 15 + 6) / 7 + 256 + 64 + 15) & ~15ull);
}

/* HEADER callback -> host -> here */
int init_decompressors(gpup_header_info* h, gpup_image* image)
{
  DecodeCtx* C = g_ctx;
  if(!C || !h || !image || !image->comps)
    return -1;
  b2k_coding& cp = C->cp;
  memset(&cp, 0, sizeof(cp));
  C->rc = 1; /* "not handled" unless everything below fits */
  if(!(h->cblk_sty & GPUP_CBLKSTY_HT) || h->t_grid_width != 1 || h->t_grid_height != 1 || h->mct > 1)
    return 0;
  cp.x0 = image->x0; cp.y0 = image->y0; cp.x1 = image->x1; cp.y1 = image->y1;
  cp.numcomps = image->numcomps;
  if(cp.numcomps < 1 || cp.numcomps > 4)
    return 0;
  cp.prec = image->comps[0].prec;
  cp.sgnd = image->comps[0].sgnd;
  for(uint16_t c = 0; c < image->numcomps; ++c)
  {
    const gpup_image_comp& k = image->comps[c];
    if(k.dx != 1 || k.dy != 1 || k.prec != cp.prec || (k.sgnd ? 1 : 0) != cp.sgnd)
      return 0;
  }
  cp.numres = h->numresolutions;
  cp.cblkw_exp = (uint8_t)floor_log2_u32(h->cblockw_init);
  cp.cblkh_exp = (uint8_t)floor_log2_u32(h->cblockh_init);
  cp.irreversible = h->irreversible;
  cp.mct = h->mct;
  cp.numgbits = 1; /* Grok's HT setting (GrkCompress.cpp L849); checked per block against numBitPlanes */
  for(int r = 0; r < 33; ++r)
  {
    cp.prcw_exp[r] = r < h->numresolutions && h->prcw_init[r] ? (uint8_t)floor_log2_u32(h->prcw_init[r]) : 15;
    cp.prch_exp[r] = r < h->numresolutions && h->prch_init[r] ? (uint8_t)floor_log2_u32(h->prch_init[r]) : 15;
    /* the host hands 1 << PPx; PPx = 0 would read as "default" in b2k_coding -> leave such streams to the host */
    if(r < h->numresolutions && (h->prcw_init[r] == 1 || h->prch_init[r] == 1))
      return 0;
  }
  if(unsupported_reason(cp))
    return 0;
  /* enumerate, allocate the block buffers and the tree */
  const int64_t n = b2k_enumerate(&cp, 1, 0, nullptr, 0);
  if(n < 0)
    return 0;
  C->blocks.resize((size_t)n);
  b2k_enumerate(&cp, 1, 0, C->blocks.data(), (uint64_t)n);
  uint64_t off = 0;
  for(b2k_block& b : C->blocks)
  {
    b.offset = off;
    b.length = (b.x1 > b.x0 && b.y1 > b.y0) ? block_capacity(b.x1 - b.x0, b.y1 - b.y0, b.kmax) : 0;
    b.numbps = 0;
    b.numpasses = 0;
    off += b.length;
  }
  C->slab_bytes = off;
  C->slab = (uint8_t*)malloc(off + 64); /* pages are touched only where the host writes */
  if(!C->slab)
    return -1;
  b2k_result fake{};
  fake.num_blocks = (uint64_t)n;
  fake.blocks = C->blocks.data();
  fake.bytes = C->slab;
  fake.num_bytes = off;
  C->tree = b2k_result_to_gpup_tile(&cp, &fake, 0);
  if(!C->tree)
    return -1;
  /* remember the node of every block (same walk order as the builder), reset what T2 will fill */
  C->cb_of.clear();
  C->bands.clear();
  for(size_t c = 0; c < C->tree->numComponents; ++c)
  {
    gpup_tile_component* tc = C->tree->tileComponents[c];
    for(size_t r = 0; r < tc->numResolutions; ++r)
      for(size_t bi = 0; bi < tc->resolutions[r]->numBands; ++bi)
      {
        gpup_band* band = tc->resolutions[r]->band[bi];
        C->bands.push_back(band);
        for(uint64_t p = 0; p < band->numPrecincts; ++p)
          for(uint64_t k = 0; k < band->precincts[p]->numBlocks; ++k)
          {
            gpup_code_block* cb = band->precincts[p]->blocks[k];
            cb->compressedDataLength = 0;
            cb->numBitPlanes = 0;
            cb->numPasses = 0;
            C->cb_of.push_back(cb);
          }
      }
  }
  if(C->cb_of.size() != C->blocks.size())
    return -1;
  C->ready = true;
  C->rc = 0;
  return 0;
}

} // namespace

extern "C" int32_t plugin_decompress(gpup_decompress_params* params, PLUGIN_DECODE_USER_CALLBACK cb)
{
  if(!cb)
    return -1;
  b2k_engine* eng = b2k_plugin_engine();
  if(!eng)
    return -1;
  DecodeCtx ctx;
  g_ctx = &ctx;
  PluginDecodeCallbackInfo info("", "", params, 1 /* J2K */, GPUP_DECODE_HEADER);
  info.init_decompressors_func = init_decompressors;
  info.deviceId = 0;
  int32_t rc = -1;
  int32_t** planes = nullptr;
  gpup_image out_img{};
  std::vector<gpup_image_comp> comps;
  std::vector<int32_t*> plane_ptrs;
  std::vector<uint32_t> strides;
  do
  {
    if(cb(&info) != 0 || !ctx.ready)
    {
      rc = ctx.rc ? ctx.rc : 1;
      break;
    }
    /* ---- T2: the host fills the tree ---- */
    info.tile = ctx.tree;
    info.decompress_flags = GPUP_DECODE_T2;
    ctx.tree->decompress_flags = GPUP_DECODE_T2;
    if(cb(&info) != 0)
    {
      rc = -1;
      break;
    }
    /* ---- gather what the host parsed; compact the used bytes for the upload ---- */
    const b2k_coding& cp = ctx.cp;
    const std::vector<BandQuant> q = band_quant(cp);
    uint64_t used = 0;
    bool ok = true;
    for(size_t i = 0; i < ctx.blocks.size() && ok; ++i)
    {
      const gpup_code_block* cbk = ctx.cb_of[i];
      b2k_block& b = ctx.blocks[i];
      const uint32_t cap = b.length;
      b.length = cbk->compressedDataLength;
      b.numbps = cbk->numBitPlanes;
      b.numpasses = (uint8_t)(cbk->numPasses > 255 ? 255 : cbk->numPasses);
      if(b.length > cap || (b.length && (b.numpasses != 1 || b.numbps > b.kmax || b.numbps < 1)))
        ok = false; /* refinement passes / other guard bits / overflow: leave it to the CPU */
      used += b.length;
    }
    if(ok && cp.irreversible)
    { /* host hands decoder-convention step / 2 (TileProcessor.cpp L183-184): it must be this engine's */
      size_t bi = 0;
      for(int c = 0; c < cp.numcomps && ok; ++c)
        for(int r = 0; r < cp.numres && ok; ++r)
          for(int b = 0; b < (r ? 3 : 1) && ok; ++b, ++bi)
          {
            const float mine = q[band_quant_index(r, r ? b + 1 : 0)].step_dec;
            const float theirs = ctx.bands[bi]->stepsize * 2.0f;
            if(std::fabs(mine - theirs) > 1e-6f * std::fabs(mine))
              ok = false;
          }
    }
    if(!ok)
    {
      rc = 1;
      break;
    }
    uint8_t* compact = (uint8_t*)b2k_host_alloc(used + 64);
    if(!compact)
    {
      rc = -1;
      break;
    }
    uint64_t at = 0;
    for(b2k_block& b : ctx.blocks)
    {
      if(b.length)
        memcpy(compact + at, ctx.slab + b.offset, b.length);
      b.offset = at;
      at += b.length;
    }
    /* ---- decode into pinned planes ---- */
    const uint32_t w = cp.x1 - cp.x0, hgt = cp.y1 - cp.y0;
    const uint32_t stride = (w + 15u) & ~15u; /* 64-byte aligned rows (gpu_plugin_shared.h L540-544) */
    comps.resize(cp.numcomps);
    plane_ptrs.resize(cp.numcomps);
    strides.assign(cp.numcomps, stride);
    bool alloc_ok = true;
    for(int c = 0; c < cp.numcomps; ++c)
    {
      plane_ptrs[c] = (int32_t*)b2k_host_alloc((size_t)stride * hgt * sizeof(int32_t));
      alloc_ok = alloc_ok && plane_ptrs[c];
    }
    double ms = 0;
    int32_t drc = alloc_ok ? b2k_decode(eng, &cp, ctx.blocks.data(), ctx.blocks.size(), compact, used, plane_ptrs.data(),
                                        strides.data(), 1, 0, &ms)
                           : -1;
    b2k_host_free(compact);
    if(drc != 0)
    {
      rc = drc > 0 ? 1 : -1;
      for(int32_t* p : plane_ptrs)
        b2k_host_free(p);
      plane_ptrs.clear();
      break;
    }
    /* ---- POST_T1: hand the planes over ---- */
    for(int c = 0; c < cp.numcomps; ++c)
    {
      gpup_image_comp& k = comps[c];
      memset(&k, 0, sizeof(k));
      k.x0 = cp.x0; k.y0 = cp.y0; k.w = w; k.h = hgt; k.stride = stride;
      k.dx = k.dy = 1;
      k.prec = cp.prec;
      k.sgnd = cp.sgnd;
      k.data = plane_ptrs[c];
      k.owns_data = false;
    }
    out_img.x0 = cp.x0; out_img.y0 = cp.y0; out_img.x1 = cp.x1; out_img.y1 = cp.y1;
    out_img.numcomps = cp.numcomps;
    out_img.color_space = info.image ? info.image->color_space : 0;
    out_img.comps = comps.data();
    gpup_image* host_img = info.image;
    info.image = &out_img;
    info.plugin_owns_image = true;
    info.decompress_flags = GPUP_DECODE_POST_T1;
    const int32_t prc = cb(&info);
    info.image = host_img;
    for(int32_t* p : plane_ptrs)
      b2k_host_free(p);
    plane_ptrs.clear();
    rc = prc == 0 ? 0 : -1;
  } while(false);
  (void)planes;
  /* ---- CLEAN ---- */
  info.decompress_flags = GPUP_DECODE_CLEAN;
  cb(&info);
  if(ctx.tree)
    b2k_plugin_free_tree(ctx.tree);
  free(ctx.slab);
  g_ctx = nullptr;
  return rc;
}
