/* Plain-C use of the boundary (include/grok_b200.h): encode a synthetic 12-bit RGB image, write an HTJ2K codestream,
 * parse it again and decode it in place from the file bytes.
 *   gcc -O2 -Iinclude examples/roundtrip.c -Lgrok_b200 -lgrokj2k_plugin -Wl,-rpath,$PWD/grok_b200 -o /tmp/roundtrip && /tmp/roundtrip out.j2c */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "grok_b200.h"

int main(int argc, char** argv)
{
  const uint32_t W = 1920, H = 1080, NC = 3;
  b2k_engine* eng = NULL;
  if(b2k_engine_create(0, &eng) != 0)
  {
    fprintf(stderr, "engine: %s\n", b2k_last_error()); /* no CUDA device: the engine has no CPU path */
    return 2;
  }
  b2k_coding cp;
  memset(&cp, 0, sizeof(cp));
  cp.x1 = W; cp.y1 = H;
  cp.tw = 1024; cp.th = 1024;                    /* 2 x 2 tiles (the last row / column ragged) */
  cp.numcomps = NC; cp.prec = 12; cp.numres = 6; /* 5 wavelet levels */
  cp.cblkw_exp = cp.cblkh_exp = 6;               /* 64 x 64 code blocks */
  cp.mct = 1; cp.numgbits = 1;
  for(int r = 0; r < 33; ++r) cp.prcw_exp[r] = cp.prch_exp[r] = 15;

  int32_t *in[3], *out[3];
  uint32_t strides[3] = {W, W, W};
  for(uint32_t c = 0; c < NC; ++c)
  {
    in[c] = (int32_t*)b2k_host_alloc((size_t)W * H * 4); /* pinned */
    out[c] = (int32_t*)b2k_host_alloc((size_t)W * H * 4);
    for(uint32_t y = 0; y < H; ++y)
      for(uint32_t x = 0; x < W; ++x)
        in[c][(size_t)y * W + x] = (int32_t)((x * (3 + c) + y * (5 - c)) / 16 + ((x * 2654435761u + y * 40503u) >> 27)) & 4095;
  }

  b2k_result* res = NULL;
  if(b2k_encode(eng, &cp, (const int32_t* const*)in, strides, 1, 0, &res) != 0)
  {
    fprintf(stderr, "encode: %s\n", b2k_last_error());
    return 1;
  }
  const int64_t n = b2k_codestream_write(&cp, res, B2K_CS_TLM | B2K_CS_PLT, NULL, 0);
  uint8_t* cs = (uint8_t*)b2k_host_alloc((size_t)n);
  if(n < 0 || b2k_codestream_write(&cp, res, B2K_CS_TLM | B2K_CS_PLT, cs, (uint64_t)n) != n)
  {
    fprintf(stderr, "codestream: %s\n", b2k_last_error());
    return 1;
  }
  printf("%u blocks, %llu coded bytes, codestream %lld bytes (%.2f bpp), encode %.2f ms on the device\n", (unsigned)res->num_blocks,
         (unsigned long long)res->num_bytes, (long long)n, 8.0 * (double)n / ((double)W * H), res->ms_total);
  b2k_result_free(res);
  if(argc > 1)
  {
    FILE* f = fopen(argv[1], "wb");
    if(f) { fwrite(cs, 1, (size_t)n, f); fclose(f); }
  }

  b2k_coding cp2;
  const int64_t nb = b2k_codestream_parse(cs, (uint64_t)n, &cp2, NULL, 0);
  b2k_block* blocks = nb > 1 ? (b2k_block*)malloc(sizeof(b2k_block) * (size_t)nb) : NULL;
  double ms = 0;
  if(nb <= 1 || b2k_codestream_parse(cs, (uint64_t)n, &cp2, blocks, (uint64_t)nb) != nb ||
     b2k_decode(eng, &cp2, blocks, (uint64_t)nb, cs, (uint64_t)n, out, strides, 1, 0, &ms) != 0)
  {
    fprintf(stderr, "decode: %s\n", b2k_last_error());
    return 1;
  }
  size_t bad = 0;
  for(uint32_t c = 0; c < NC; ++c)
    for(size_t i = 0; i < (size_t)W * H; ++i)
      bad += in[c][i] != out[c][i];
  printf("decode %.2f ms, %zu samples differ (lossless: 0)\n", ms, bad);
  free(blocks);
  b2k_host_free(cs);
  for(uint32_t c = 0; c < NC; ++c) { b2k_host_free(in[c]); b2k_host_free(out[c]); }
  b2k_engine_destroy(eng);
  return bad != 0;
}
