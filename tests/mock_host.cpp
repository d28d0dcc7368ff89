/* tests/mock_host.cpp -- TEST INFRASTRUCTURE: a stand-in for Grok's side of the plugin_decompress()
 * callback protocol (CodeStreamDecompress.cpp L217-271, TileProcessor.cpp L157-227), fed with a block
 * table + byte arena instead of a parsed codestream.  It does what the host does in each phase:
 * HEADER: fill header_info + an image shell, call the plugin's init_decompressors_func;
 * T2: walk the plugin-allocated tree in Grok's order and copy every block's bytes / numbps / passes
 *     into it, set band stepsize/2; POST_T1: copy the decoded planes out; CLEAN: nothing. */
#include <cstdint>
#include <cstring>
#include <vector>
#include "plugin_decode_abi.h"

struct MockState
{
  const b2k_coding* cp;
  const b2k_block* blocks;
  uint64_t nblocks;
  const uint8_t* bytes;
  const float* band_step_dec; /* comp-major, resolution, band */
  int32_t** out;
  const uint32_t* out_strides;
  gpup_image img;
  std::vector<gpup_image_comp> comps;
  int phases;
  int tamper; /* 1: claim two passes on the first coded block (must come back "not handled") */
};
static MockState* S = nullptr;

static int32_t host_callback(PluginDecodeCallbackInfo* info)
{
  const b2k_coding& cp = *S->cp;
  if(info->decompress_flags & GPUP_DECODE_CLEAN)
  {
    S->phases |= 8;
    return 0;
  }
  if(info->decompress_flags & GPUP_DECODE_HEADER)
  {
    S->phases |= 1;
    gpup_header_info& h = info->header_info;
    memset(&h, 0, sizeof(h));
    h.cblockw_init = 1u << cp.cblkw_exp;
    h.cblockh_init = 1u << cp.cblkh_exp;
    h.irreversible = cp.irreversible;
    h.mct = cp.mct;
    h.numresolutions = cp.numres;
    h.cblk_sty = GPUP_CBLKSTY_HT;
    for(int r = 0; r < cp.numres; ++r)
    {
      h.prcw_init[r] = 1u << cp.prcw_exp[r];
      h.prch_init[r] = 1u << cp.prch_exp[r];
    }
    h.tx0 = cp.x0; h.ty0 = cp.y0; h.t_width = cp.x1 - cp.x0; h.t_height = cp.y1 - cp.y0;
    h.t_grid_width = h.t_grid_height = 1;
    h.max_layers_ = 1;
    S->comps.assign(cp.numcomps, gpup_image_comp{});
    for(int c = 0; c < cp.numcomps; ++c)
    {
      gpup_image_comp& k = S->comps[c];
      k.x0 = cp.x0; k.y0 = cp.y0; k.w = cp.x1 - cp.x0; k.h = cp.y1 - cp.y0; k.stride = k.w;
      k.dx = k.dy = 1; k.prec = cp.prec; k.sgnd = cp.sgnd; k.data = nullptr; k.owns_data = false;
    }
    S->img.x0 = cp.x0; S->img.y0 = cp.y0; S->img.x1 = cp.x1; S->img.y1 = cp.y1;
    S->img.numcomps = cp.numcomps; S->img.color_space = 2; S->img.comps = S->comps.data();
    info->image = &S->img;
    if(!info->init_decompressors_func)
      return -1;
    return info->init_decompressors_func(&info->header_info, info->image);
  }
  if(info->decompress_flags & GPUP_DECODE_T2)
  {
    S->phases |= 2;
    gpup_tile* T = info->tile;
    if(!T || T->numComponents != cp.numcomps)
      return -1;
    uint64_t idx = 0, band_idx = 0;
    bool tampered = false;
    for(size_t c = 0; c < T->numComponents; ++c)
    {
      gpup_tile_component* tc = T->tileComponents[c];
      if(tc->numResolutions != cp.numres)
        return -1;
      for(size_t r = 0; r < tc->numResolutions; ++r)
        for(size_t b = 0; b < tc->resolutions[r]->numBands; ++b, ++band_idx)
        {
          gpup_band* band = tc->resolutions[r]->band[b];
          band->stepsize = S->band_step_dec[band_idx] / 2; /* TileProcessor.cpp L183-184 */
          for(uint64_t p = 0; p < band->numPrecincts; ++p)
            for(uint64_t k = 0; k < band->precincts[p]->numBlocks; ++k, ++idx)
            {
              if(idx >= S->nblocks)
                return -1;
              const b2k_block& s = S->blocks[idx];
              gpup_code_block* cb = band->precincts[p]->blocks[k];
              if(cb->x0 != s.x0 || cb->y0 != s.y0 || cb->x1 != s.x1 || cb->y1 != s.y1 || s.precno != p || s.cblkno != k)
                return -2; /* tree order differs from the enumeration */
              if(!s.length)
                continue;
              memcpy(cb->compressedData, S->bytes + s.offset, s.length);
              cb->compressedDataLength = s.length;
              cb->numBitPlanes = s.numbps;
              cb->numPasses = s.numpasses;
              if(S->tamper == 1 && !tampered)
              {
                cb->numPasses = 2;
                tampered = true;
              }
            }
        }
    }
    return idx == S->nblocks ? 0 : -1;
  }
  if(info->decompress_flags & GPUP_DECODE_POST_T1)
  {
    S->phases |= 4;
    const gpup_image* im = info->image;
    if(!im || im->numcomps != cp.numcomps)
      return -1;
    for(int c = 0; c < cp.numcomps; ++c)
    {
      const gpup_image_comp& k = im->comps[c];
      const uint32_t st = k.stride ? k.stride : k.w;
      for(uint32_t y = 0; y < k.h; ++y)
        memcpy(S->out[c] + (size_t)y * S->out_strides[c], k.data + (size_t)y * st, (size_t)k.w * sizeof(int32_t));
    }
    return 0;
  }
  return -1;
}

extern "C" __attribute__((visibility("default"))) int mock_host_run(void* plugin_decompress_fn, const b2k_coding* cp,
                                                                    const b2k_block* blocks, uint64_t nblocks,
                                                                    const uint8_t* bytes, const float* band_step_dec,
                                                                    int32_t** out, const uint32_t* out_strides, int tamper,
                                                                    int* phases)
{
  typedef int32_t (*FN)(gpup_decompress_params*, PLUGIN_DECODE_USER_CALLBACK);
  MockState st{};
  st.cp = cp; st.blocks = blocks; st.nblocks = nblocks; st.bytes = bytes; st.band_step_dec = band_step_dec;
  st.out = out; st.out_strides = out_strides; st.tamper = tamper;
  S = &st;
  gpup_decompress_params params;
  memset(&params, 0, sizeof(params));
  const int32_t rc = ((FN)plugin_decompress_fn)(&params, host_callback);
  if(phases)
    *phases = st.phases;
  S = nullptr;
  return rc;
}
