/*
 * include/grok_b200.h -- C ABI of libgrokj2k_plugin.so, the B200-native JPEG 2000 tile engine.
 *
 * Two groups of entry points:
 *
 *  (1) The STOCK accelerator-plugin symbols Grok's host library resolves with dlsym()
 *      (reference: src/lib/core/grok.cpp L1177-1186, L1297-1300; typedefs
 *      src/lib/core/plugin/plugin_interface.h L50-133; structs
 *      src/lib/core/plugin/gpup/gpu_plugin_shared.h L215-530).  The struct layouts below are
 *      binary-compatible restatements of that contract: field order and types must not change.
 *
 *  (2) The tile-aware b2k_* engine API.  The stock contract is "whole image = one tile"
 *      (CodeStreamCompress.cpp L908-912, CodeStreamDecompress.cpp L199-205); multi-tile
 *      codestreams, multi-GPU sharding, device-resident buffers and per-stage parity hooks go
 *      through these.  INTEGRATION.md shows the ~30-line host patch that binds them.
 *
 * All functions are extern "C", plain pointers and sizes, no C++ or torch types.
 * Return convention (plugin_accelerate.h L32-36): 0 = handled, >0 = not handled (host falls
 * back to its CPU path), <0 = device failure.
 */
#ifndef GROK_B200_H
#define GROK_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2K_API __attribute__((visibility("default")))

/* ============================================================================================
 * (1) stock plugin contract -- gpu_plugin_shared.h
 * ========================================================================================== */
#define GPUP_PATH_LEN 4096
#define GPUP_MAX_LAYERS 256
#define GPUP_MAX_DECOMP_LVLS 32
#define GPUP_MAXRLVLS (GPUP_MAX_DECOMP_LVLS + 1)
#define GPUP_MAX_SUPPORTED_PREC 16
#define GPUP_BIBO_EXTRA_BITS 7
#define GPUP_MAX_PASSES (3 * (GPUP_MAX_SUPPORTED_PREC + GPUP_BIBO_EXTRA_BITS) - 2)
#define GPUP_BUFFER_ALIGNMENT 64

#define GPUP_DECODE_HEADER (1 << 0)
#define GPUP_DECODE_T2 (1 << 1)
#define GPUP_DECODE_T1 (1 << 2)
#define GPUP_DECODE_POST_T1 (1 << 3)
#define GPUP_DECODE_CLEAN (1 << 4)

#define GPUP_STATE_NO_DEBUG 0x0
#define GPUP_CBLKSTY_HT 0x040

/* enums of gpu_plugin_shared.h L63-146 are plain C enums (int sized) */
typedef int32_t gpup_prog_order;
typedef int32_t gpup_color_space;
typedef int32_t gpup_file_fmt;
typedef int32_t gpup_codec_fmt;
typedef int32_t gpup_rate_control;

typedef struct _gpup_image_comp /* L215-225 */
{
  uint32_t x0, y0;
  uint32_t w;
  uint32_t stride;
  uint32_t h;
  uint8_t dx, dy;
  uint8_t prec;
  bool sgnd;
  int32_t* data;
  bool owns_data;
} gpup_image_comp;

typedef struct _gpup_image /* L227-233 */
{
  uint32_t x0, y0, x1, y1;
  uint16_t numcomps;
  gpup_color_space color_space;
  gpup_image_comp* comps;
} gpup_image;

typedef struct _gpup_pass /* L240-245 */
{
  double distortionDecrease;
  size_t rate;
  size_t length;
} gpup_pass;

typedef struct _gpup_code_block /* L247-261 */
{
  uint32_t x0, y0, x1, y1;
  unsigned int* contextStream;
  uint32_t numPix;
  uint8_t* compressedData;
  uint32_t compressedDataLength;
  uint8_t numBitPlanes;
  size_t numPasses;
  gpup_pass passes[GPUP_MAX_PASSES];
  unsigned int sortedIndex;
} gpup_code_block;

typedef struct _gpup_precinct
{
  uint64_t numBlocks;
  gpup_code_block** blocks;
} gpup_precinct;

typedef struct _gpup_band
{
  uint8_t orientation;
  uint64_t numPrecincts;
  gpup_precinct** precincts;
  float stepsize;
} gpup_band;

typedef struct _gpup_resolution
{
  size_t level;
  size_t numBands;
  gpup_band** band;
} gpup_resolution;

typedef struct _gpup_tile_component
{
  size_t numResolutions;
  gpup_resolution** resolutions;
} gpup_tile_component;

typedef struct _gpup_tile /* L289-294 */
{
  uint32_t decompress_flags;
  size_t numComponents;
  gpup_tile_component** tileComponents;
} gpup_tile;

typedef struct _gpup_header_info /* L300-318 */
{
  uint32_t cblockw_init;
  uint32_t cblockh_init;
  bool irreversible;
  uint8_t mct;
  uint16_t rsiz;
  uint8_t numresolutions;
  gpup_prog_order prog_order;
  uint8_t csty;
  uint8_t cblk_sty;
  uint32_t prcw_init[GPUP_MAXRLVLS];
  uint32_t prch_init[GPUP_MAXRLVLS];
  uint32_t tx0, ty0;
  uint32_t t_width, t_height;
  uint16_t t_grid_width, t_grid_height;
  uint16_t max_layers_;
} gpup_header_info;

typedef struct _gpup_compress_params /* L324-374 */
{
  bool tile_size_on;
  uint32_t tx0, ty0, t_width, t_height;
  uint16_t numlayers;
  bool allocationByRateDistoration;
  double layer_rate[GPUP_MAX_LAYERS];
  bool allocationByQuality;
  double layer_distortion[GPUP_MAX_LAYERS];
  uint8_t csty;
  uint8_t numgbits;
  gpup_prog_order prog_order;
  uint32_t numpocs;
  uint8_t numresolution;
  uint32_t cblockw_init;
  uint32_t cblockh_init;
  uint8_t cblk_sty;
  bool irreversible;
  int32_t roi_compno;
  uint32_t roi_shift;
  uint32_t res_spec;
  uint32_t prcw_init[GPUP_MAXRLVLS];
  uint32_t prch_init[GPUP_MAXRLVLS];
  char infile[GPUP_PATH_LEN];
  char outfile[GPUP_PATH_LEN];
  uint32_t image_offset_x0;
  uint32_t image_offset_y0;
  uint8_t subsampling_dx;
  uint8_t subsampling_dy;
  gpup_file_fmt decod_format;
  gpup_file_fmt cod_format;
  bool enableTilePartGeneration;
  uint8_t newTilePartProgressionDivider;
  uint8_t mct;
  uint64_t max_cs_size;
  uint64_t max_comp_size;
  uint16_t rsiz;
  uint16_t framerate;
  gpup_rate_control rateControlAlgorithm;
  uint32_t numThreads;
  int32_t deviceId;
  uint32_t duration;
  uint32_t kernelBuildOptions;
  uint32_t repeats;
  bool verbose;
  bool sharedMemoryInterface;
  bool apply_xyz_transform;
} gpup_compress_params;

typedef struct _gpup_decompress_core_params
{
  uint8_t reduce;
  uint16_t layers_to_decompress_;
} gpup_decompress_core_params;

typedef struct _gpup_decompress_params /* L386-401 */
{
  gpup_decompress_core_params core;
  char infile[GPUP_PATH_LEN];
  char outfile[GPUP_PATH_LEN];
  gpup_codec_fmt decod_format;
  gpup_file_fmt cod_format;
  double dw_x0, dw_y0, dw_x1, dw_y1;
  uint16_t tileIndex;
  int32_t deviceId;
  uint32_t kernelBuildOptions;
  uint32_t repeats;
  uint32_t numThreads;
  bool verbose_;
  void* user_data;
} gpup_decompress_params;

typedef struct _gpup_init_info /* L403-409 */
{
  int32_t deviceId;
  bool verbose;
  const char* license;
  const char* server;
} gpup_init_info;

typedef int (*GPUP_INIT_DECOMPRESSORS)(gpup_header_info* header_info, gpup_image* image);

typedef struct _gpup_decompress_callback_info /* L503-524 */
{
  size_t deviceId;
  GPUP_INIT_DECOMPRESSORS init_decompressors_func;
  const char* input_file_name;
  const char* output_file_name;
  gpup_codec_fmt decod_format;
  gpup_file_fmt cod_format;
  void* codec;
  gpup_header_info header_info;
  gpup_decompress_params* decompressor_parameters;
  gpup_image* image;
  bool plugin_owns_image;
  gpup_tile* tile;
  unsigned int error_code;
  uint32_t decompress_flags;
  uint32_t full_image_x0;
  uint32_t full_image_y0;
  void* user_data;
  void* format_private;
} gpup_decompress_callback_info;

typedef int32_t (*GPUP_DECOMPRESS_USER_CALLBACK)(gpup_decompress_callback_info* info);

/* minpf loader handshake: minpf_plugin.h L98-111 */
typedef int32_t (*minpf_exit_func)(void);
typedef struct _minpf_platform_services minpf_platform_services; /* opaque here; see INTEGRATION.md */

/* --- symbols resolved by name by libgrokj2k (grok.cpp L1177-1186, L1297-1298) --- */
B2K_API minpf_exit_func minpf_post_load_plugin(const minpf_platform_services* services); /* minpf_plugin.h L109 */
B2K_API bool plugin_init(gpup_init_info info);                    /* plugin_interface.h L60 */
B2K_API uint32_t plugin_get_debug_state(void);                    /* plugin_interface.h L56 */
B2K_API int32_t gpup_encode_mem(gpup_compress_params* params, gpup_image* image,
                                gpup_tile** out);                 /* grok.cpp L1299-1300 */
B2K_API void gpup_tile_free(gpup_tile* tile);                     /* grok.cpp L1330 */
/* --- optional symbols the PATCHED host resolves (baseline/patches/0001-multi-tile-plugin-encode-decode.patch;
 * the stock contract is one tile per image, CodeStreamCompress.cpp L908-912 / CodeStreamDecompress.cpp L199-205) --- */
/* all tiles of a multi-tile image in one call; (*out_tiles)[t] is the stock tree of tile index t in the tile's
 * canvas coordinates; the array and the trees stay valid until gpup_tiles_free(*out_tiles, n) */
B2K_API int32_t gpup_encode_mem_tiles(gpup_compress_params* params, gpup_image* image, gpup_tile*** out_tiles,
                                      uint32_t* out_num_tiles);
B2K_API void gpup_tiles_free(gpup_tile** tiles, uint32_t num_tiles);
/* a whole (multi-tile) code stream the host holds in memory -> the int32 planes of `image` (allocated by the host) */
B2K_API int32_t plugin_decompress_codestream(const uint8_t* codestream, uint64_t length, gpup_image* image);

/* --- in-memory batch compress (grok.cpp L1538-1545, L1655-1857): frames of one shape, several in flight --- */
typedef struct _gpup_stream_params /* gpu_plugin_shared.h L415-421 */
{
  const char* file;
  uint8_t* buf;
  size_t buf_len;
  size_t buf_compressed_len;
} gpup_stream_params;
typedef struct _gpup_compress_callback_info /* L428-440 */
{
  const char* input_file_name;
  bool outputFileNameIsRelative;
  const char* output_file_name;
  gpup_compress_params* compressor_parameters;
  gpup_image* image;
  gpup_tile* tile;
  gpup_stream_params stream_params;
  unsigned int error_code;
  void* host_data;
} gpup_compress_callback_info;
typedef uint64_t (*GPUP_COMPRESS_USER_CALLBACK)(gpup_compress_callback_info* info); /* L442 */
typedef enum { GPUP_SOURCE_PLANAR_RGB = 0, GPUP_SOURCE_YUV420P = 1, GPUP_SOURCE_YUV422P = 2, GPUP_SOURCE_RGB48LE = 3 } gpup_source_format; /* L127-136 */
typedef enum { GPUP_YUV_BT601 = 0, GPUP_YUV_BT709 = 1, GPUP_YUV_BT2020 = 2 } gpup_yuv_matrix;                                             /* L139-144 */
typedef struct _gpup_batch_memory_info /* L454-473 */
{
  gpup_compress_params* compress_parameters;
  uint32_t width, height, numcomps;
  uint32_t source_prec; /* bits per sample the caller submits */
  uint32_t prec;        /* bits per sample the code stream carries */
  GPUP_COMPRESS_USER_CALLBACK callback;
  bool xyz_on_device;   /* written by begin */
  gpup_source_format source_format;
  gpup_yuv_matrix yuv_matrix;
  bool yuv_full_range;
} gpup_batch_memory_info;
B2K_API int32_t gpup_batch_memory_begin(gpup_batch_memory_info* info);
B2K_API bool gpup_batch_memory_submit(const uint8_t* packed, void* host_data);
B2K_API bool gpup_batch_memory_submit_planes(const uint8_t* const planes[3], const size_t stride_bytes[3], void* host_data);
B2K_API bool gpup_batch_memory_end(void);
/* --- in-memory batch decompress (grok.cpp L2023-2188): the plugin's workers pull code streams of one shape --- */
typedef bool (*GPUP_BATCH_DECOMPRESS_PULL)(void* user, const uint8_t** codestream, size_t* length, void** frame_user); /* L477-478 */
typedef struct _gpup_display_transform /* L481-487 */
{
  const float* transfer;
  const float* matrix;
} gpup_display_transform;
typedef struct _gpup_batch_decompress_memory_info /* L492-506 */
{
  gpup_decompress_params* decompress_parameters;
  gpup_header_info header_info;
  gpup_image* image;
  GPUP_BATCH_DECOMPRESS_PULL pull;
  void* pull_user;
  bool srgb8_output;
  const gpup_display_transform* display_transform;
  bool rgb8_on_device; /* written by begin: stays false here (frames come back as int32 planes) */
} gpup_batch_decompress_memory_info;
/* plugin_batch_decompress_memory_begin(info, PLUGIN_DECODE_USER_CALLBACK) takes the same C++ callback type as
 * plugin_decompress (below), so it is declared with it; its end has a plain C signature: */
B2K_API bool plugin_batch_decompress_memory_end(void);            /* plugin_interface.h L133 */
/* plugin_decompress (plugin_interface.h L117-120) takes a C++ struct with std::string members
 * (PluginDecodeCallbackInfo L78-115): it is declared in grok_b200/csrc/plugin_decode.cpp and
 * documented in INTEGRATION.md, not here, so that this header stays C. */

/* ============================================================================================
 * (2) tile-aware engine API
 * ========================================================================================== */
typedef struct b2k_engine b2k_engine;

/* coding parameters of one image; the subset of grk_cparameters / SIZ+COD+QCD the tile engine
 * depends on (CodeStreamCompress::init, CodeStreamCompress.cpp L229-855) */
typedef struct b2k_coding
{
  uint32_t x0, y0, x1, y1;         /* image area on the canvas (SIZ) */
  uint32_t tx0, ty0, tw, th;       /* tile grid origin and nominal tile size (tw==0: one tile) */
  uint16_t numcomps;               /* 1..4; all components dx=dy=1 and same precision */
  uint8_t prec;                    /* bits per sample */
  uint8_t sgnd;                    /* 1 = signed samples */
  uint8_t numres;                  /* resolutions = decomposition levels + 1 */
  uint8_t cblkw_exp, cblkh_exp;    /* log2 nominal code-block size (6,6 = 64x64) */
  uint8_t irreversible;            /* 0: 5/3 + RCT, 1: 9/7 + ICT */
  uint8_t mct;                     /* 1: colour transform on components 0..2 */
  uint8_t numgbits;                /* guard bits; Grok's HT CLI forces 1 (GrkCompress.cpp L849) */
  uint8_t prcw_exp[33], prch_exp[33]; /* precinct exponents per resolution, 1..15 (15 = maximal); 0 reads as 15.
                                         A true exponent 0 (1-sample precincts) is declined by every entry point */
  uint8_t cblk_sty;                /* code-block style bits (COD); only 0x08 = vertically stripe-causal matters,
                                      and only to the decoder's SigProp pass (CoderOJPH.cpp L248) */
  uint8_t qcd_explicit;            /* 0: band exponents / mantissas are the HT quantiser's (QuantizerOJPH.cpp L193-259),
                                      as Grok's encoder signals them.  1: take them from qcd_expn / qcd_mant below --
                                      what b2k_codestream_parse fills in for a foreign stream's QCD (band order of QCD:
                                      LL, then HL, LH, HH per resolution) */
  uint8_t qcd_expn[97];
  uint16_t qcd_mant[97];
} b2k_coding;

/* One coded block as the host's T2 needs it (cf. compress_synch_with_plugin,
 * plugin_bridge.cpp L113-243), in Grok's enumeration order tile->comp->res->band->prec->cblk */
typedef struct b2k_block
{
  uint32_t tile;                   /* tile index, raster order */
  uint16_t comp;
  uint8_t resno, band_index, orient;
  uint8_t kmax;                    /* band->maxBitPlanes_ (TileProcessor.cpp L417-419) */
  uint8_t numbps;                  /* coded bit planes as T2 signals them: Kmax - zero bit planes.
                                      The encoder returns 1 (CoderOJPH.cpp L203-206) */
  uint8_t numpasses;               /* 0 not coded; 1 HT cleanup (all the encoder emits); decode also takes 2
                                      (+ SigProp) and 3 (+ SigProp + MagRef) from foreign streams */
  uint32_t precno, cblkno;
  uint32_t x0, y0, x1, y1;         /* block rect, band canvas coordinates */
  uint32_t buf_x, buf_y;           /* position inside the tile-component Mallat buffer */
  uint32_t length;                 /* bytes of the HT cleanup segment */
  uint64_t offset;                 /* byte offset into the result's byte arena */
  float stepsize;                  /* band step size (encoder convention) */
  uint32_t length2;                /* decode: bytes of the refinement segment (SigProp, MagRef) that follows
                                      the cleanup segment at offset + length; 0 from the encoder */
} b2k_block;

typedef struct b2k_result
{
  uint64_t num_blocks;
  b2k_block* blocks;               /* host memory, owned by the result */
  uint8_t* bytes;                  /* host memory (pinned), owned by the result */
  uint64_t num_bytes;
  uint32_t num_tiles;
  double ms_h2d, ms_dwt, ms_t1, ms_d2h, ms_total; /* device-event timings of the call */
} b2k_result;

B2K_API int32_t b2k_engine_create(int32_t device, b2k_engine** out);
B2K_API void b2k_engine_destroy(b2k_engine* e);
B2K_API const char* b2k_last_error(void);

/* the b2k_coding that gpup_encode_mem (allow_tiles = 0) / gpup_encode_mem_tiles (1) derive from the host's stock
 * parameters, precinct sizes as CodeStreamCompress.cpp L793-825 derives them (0 handled, 1 not handled) */
B2K_API int32_t b2k_coding_from_gpup(const gpup_compress_params* params, const gpup_image* image, int32_t allow_tiles,
                                     b2k_coding* out);

/* pinned host memory for image planes / codestream arenas (what Grok's allocator should hand
 * to grk_image when the plugin is loaded) */
B2K_API void* b2k_host_alloc(size_t bytes);
B2K_API void b2k_host_free(void* p);
/* b2k_encode / b2k_decode can carry samples of <= 16 bits over PCIe in 16-bit containers: each
 * pipeline chunk is narrowed (widened) between the caller's int32 planes and a pinned staging
 * buffer by host threads while its neighbour is on the bus.  That halves the PCIe bytes and costs
 * host DRAM bandwidth, so whether it pays depends on the machine and on what else runs on it.
 *   n < 0 (default): min(cores the process may run on, 24) threads (env B2K_HOST_THREADS); each
 *          job times both ways on its first calls and keeps the faster one;
 *   n = 0: never (int32 planes are copied as they are and should be pinned);
 *   n > 0: always, with n threads (also the way to feed unpinned planes).
 * Returns the thread count in force.  b2k_host_pack_last(decode) tells which way the last
 * b2k_encode (0) / b2k_decode (1) went: 1 packed, 0 direct, -1 none yet. */
B2K_API int32_t b2k_set_host_threads(int32_t n);
B2K_API int32_t b2k_host_pack_last(int32_t decode);

/* Encode every tile of the image for which (tile_index % tile_mod) == tile_rem (tile_mod=1:
 * all tiles).  planes[c] = int32 samples, row stride strides[c] elements, origin (x0,y0). */
B2K_API int32_t b2k_encode(b2k_engine* e, const b2k_coding* cp, const int32_t* const* planes,
                           const uint32_t* strides, uint32_t tile_mod, uint32_t tile_rem,
                           b2k_result** out);
/* same, 16-bit unsigned/signed sample containers (cf. gpup_batch_memory_submit_planes) */
B2K_API int32_t b2k_encode16(b2k_engine* e, const b2k_coding* cp, const uint16_t* const* planes,
                             const uint32_t* strides, uint32_t tile_mod, uint32_t tile_rem,
                             b2k_result** out);
/* same, ONE pixel-interleaved buffer of 16-bit samples (component c of pixel x at pixels[y*stride + x*numcomps + c],
 * stride in samples >= numcomps * width): the layout of gpup_batch_memory_submit's packed frames and of
 * GPUP_SOURCE_RGB48LE (grok.cpp L1806-1836, grok.h "GRK_SOURCE_RGB48LE").  The rows cross PCIe as they are and are
 * split into planes on the device -- no host pass over the samples. */
B2K_API int32_t b2k_encode16_interleaved(b2k_engine* e, const b2k_coding* cp, const uint16_t* pixels, uint32_t stride,
                                         uint32_t tile_mod, uint32_t tile_rem, b2k_result** out);
B2K_API void b2k_result_free(b2k_result* r);

/* Decode: blocks[] (same enumeration, with length/offset filled by the host's T2 parse) and the
 * byte arena in; planes out (int32, clamped, DC shift restored). */
B2K_API int32_t b2k_decode(b2k_engine* e, const b2k_coding* cp, const b2k_block* blocks,
                           uint64_t num_blocks, const uint8_t* bytes, uint64_t num_bytes,
                           int32_t* const* planes, const uint32_t* strides, uint32_t tile_mod,
                           uint32_t tile_rem, double* ms_total);
/* b2k_decode / b2k_decode16 returning only `window` = (x0, y0, x1, y1), in cp's canvas coordinates, of the pixels:
 * planes[c] holds the window's rows (strides[c] samples apart), sample_bytes = 4 (int32) or 2 (16-bit containers).  The tiles
 * of cp are decoded as usual; only the window's pixels cross PCIe.  Made for the virtual coding of
 * b2k_codestream_parse_window (windowed / reduced decode). */
B2K_API int32_t b2k_decode_window(b2k_engine* e, const b2k_coding* cp, const b2k_block* blocks, uint64_t num_blocks,
                                  const uint8_t* bytes, uint64_t num_bytes, void* const* planes, const uint32_t* strides,
                                  const uint32_t* window, uint32_t sample_bytes, double* ms_total);
/* same, pixels returned in 16-bit containers (reversible path) */
B2K_API int32_t b2k_decode16(b2k_engine* e, const b2k_coding* cp, const b2k_block* blocks,
                             uint64_t num_blocks, const uint8_t* bytes, uint64_t num_bytes,
                             uint16_t* const* planes, const uint32_t* strides, uint32_t tile_mod,
                             uint32_t tile_rem, double* ms_total);

/* Geometry only (host): enumerate the blocks of the selected tiles, lengths zero.  Returns the
 * count; fills at most cap entries. */
B2K_API int64_t b2k_enumerate(const b2k_coding* cp, uint32_t tile_mod, uint32_t tile_rem,
                              b2k_block* out, uint64_t cap);

/* Build / free the stock gpup_tile tree for ONE tile from a result (what gpup_encode_mem
 * returns; layout rules plugin_bridge.cpp L62-111). */
B2K_API gpup_tile* b2k_result_to_gpup_tile(const b2k_coding* cp, const b2k_result* r, uint32_t tile);

/* ---- device-resident path (inputs already in HBM; what bench.py's `value` times) ---------- */
typedef struct b2k_device_job b2k_device_job;
B2K_API int32_t b2k_job_create(b2k_engine* e, const b2k_coding* cp, uint32_t tile_mod,
                               uint32_t tile_rem, b2k_device_job** out);
B2K_API void b2k_job_destroy(b2k_device_job* j);
/* upload planes into the job's device image (untimed set-up for the device-resident bench) */
B2K_API int32_t b2k_job_upload(b2k_device_job* j, const int32_t* const* planes, const uint32_t* strides);
/* run stages on device-resident data; each returns 0 and the elapsed device ms via *ms */
B2K_API int32_t b2k_job_forward(b2k_device_job* j, float* ms);  /* DC shift+MCT+DWT, all levels */
B2K_API int32_t b2k_job_t1_encode(b2k_device_job* j, float* ms, uint64_t* total_bytes);
B2K_API int32_t b2k_job_t1_decode(b2k_device_job* j, float* ms);/* from the job's own coded blocks */
B2K_API int32_t b2k_job_inverse(b2k_device_job* j, float* ms);  /* inverse DWT+MCT into image */
/* all four stages enqueued back to back, one synchronisation; stage_ms[4] optional; returns 2 once if the
 * coded size outgrew the arena (arena resized: call again) */
B2K_API int32_t b2k_job_roundtrip(b2k_device_job* j, float* ms_total, float* stage_ms, uint64_t* total_bytes);
/* `steps` round trips queued back to back with one synchronisation after the last (a benchmark loop without the
 * host in it); stage_ms[4] and level1_ms are sums over the steps, ms_total spans first start to last end. */
B2K_API int32_t b2k_job_roundtrip_n(b2k_device_job* j, uint32_t steps, float* ms_total, float* stage_ms, float* level1_ms,
                                    uint64_t* total_bytes);
/* the same round trips with the block-coder stage pipelined over `chunks` block ranges on `streams` side streams (0: the
 * defaults, 2 and 2): the transforms run alone, between them every range goes encode -> compact -> decode on its stream, so
 * the latency-bound kernels of one range run under the issue-bound kernels of its neighbours.  stage_ms[3] = forward, block
 * coder (encode + decode), inverse.  Results are those of b2k_job_roundtrip_n, byte for byte. */
B2K_API int32_t b2k_job_roundtrip_pipelined_n(b2k_device_job* j, uint32_t steps, uint32_t chunks, uint32_t streams, float* ms_total,
                                              float* stage_ms, float* level1_ms, uint64_t* total_bytes);
B2K_API int32_t b2k_job_download(b2k_device_job* j, int32_t* const* planes, const uint32_t* strides);
/* copy the coefficient planes (Mallat layout per tile, image-shaped, int32 or float bits) */
B2K_API int32_t b2k_job_download_coeffs(b2k_device_job* j, int32_t* const* planes, const uint32_t* strides);
B2K_API int32_t b2k_job_upload_coeffs(b2k_device_job* j, const int32_t* const* planes, const uint32_t* strides);
/* block-decode a caller-supplied block table + byte arena (as b2k_decode takes them) into the job's
 * coefficient planes; 1 = not handled (e.g. more than 3 HT passes), < 0 failure */
B2K_API int32_t b2k_job_t1_decode_blocks(b2k_device_job* j, const b2k_block* blocks, uint64_t num_blocks,
                                         const uint8_t* bytes, uint64_t num_bytes, float* ms);
/* fetch coded blocks of the last b2k_job_t1_encode as a host result */
B2K_API int32_t b2k_job_fetch_result(b2k_device_job* j, b2k_result** out);
B2K_API uint64_t b2k_job_num_blocks(const b2k_device_job* j);
/* Merge per-rank results (rank r coded the tiles t with t % nshards == r) into one result in full enumeration order,
 * for the writer rank after it has gathered the shards (block tables + byte arenas).  Free with b2k_result_free. */
B2K_API int32_t b2k_result_merge(const b2k_coding* cp, const b2k_result* const* shards, uint32_t nshards, b2k_result** out);

/* ---- codestream assembly / parsing on the host (the T2 step) -------------------------
 * b2k_codestream_write: a complete HTJ2K codestream (SOC, SIZ, CAP, COD, QCD, [TLM], per tile SOT [PLT] SOD
 * + packets, EOC; one layer, any of the five progression orders, optionally a tile part per resolution) from an
 * encode result that holds every tile
 * (cf. CodeStreamCompress::compress / T2Compress::compressPacket).  Returns the size; copies it to `out` if
 * cap suffices (call with out = NULL to size the buffer).  < 0 on error.
 * b2k_codestream_parse: main header -> *cp, packet headers -> block table in enumeration order with
 * numbps / numpasses / length / length2 and offsets INTO `cs`, so that b2k_decode(engine, cp, blocks, n, cs,
 * len, ...) decodes the file in place.  Returns the number of blocks (call with blocks = NULL to size the
 * table), 1 if the codestream uses something this path does not cover (b2k_last_error says what), < 0 if it
 * is damaged. */
#define B2K_CS_TLM 1u
#define B2K_CS_PLT 2u
#define B2K_CS_SOP 16u                     /* SOP marker segment before every packet */
#define B2K_CS_EPH 32u                     /* EPH marker after every packet header */
#define B2K_CS_TPARTS_R 4u                 /* one tile part per resolution (LRCP / RLCP / RPCL only), cf. grk_compress -u R */
#define B2K_CS_PROG(n) (((n) & 7u) << 8)   /* progression order: 0 LRCP (default), 1 RLCP, 2 RPCL, 3 PCRL, 4 CPRL */
B2K_API int64_t b2k_codestream_write(const b2k_coding* cp, const b2k_result* r, uint32_t flags, uint8_t* out, uint64_t cap);
B2K_API int64_t b2k_codestream_parse(const uint8_t* cs, uint64_t len, b2k_coding* cp, b2k_block* blocks, uint64_t cap_blocks);
/* Per-rank writers (tiles sharded over ranks): every rank turns ITS tiles (t % tile_mod == tile_rem, the result
 * b2k_encode gave it) into finished tile parts -- consecutive in tile order in `out`, tile_bytes[k] = length of the k-th of its
 * tiles -- and the writer rank only needs the lengths of all tiles for the header (+ TLM): code stream = header + tile parts in
 * tile-index order + 0xFFD9, byte-identical to b2k_codestream_write over the merged result.  One tile part per tile. */
B2K_API int64_t b2k_codestream_write_tiles(const b2k_coding* cp, const b2k_result* shard, uint32_t flags, uint32_t tile_mod,
                                           uint32_t tile_rem, uint8_t* out, uint64_t cap, uint64_t* tile_bytes);
B2K_API int64_t b2k_codestream_write_header(const b2k_coding* cp, uint32_t flags, const uint64_t* tile_bytes, uint32_t ntiles,
                                            uint8_t* out, uint64_t cap);
/* as b2k_codestream_write_tiles, the k-th of the shard's tiles written at out + tile_at[k] */
B2K_API int64_t b2k_codestream_write_tiles_at(const b2k_coding* cp, const b2k_result* shard, uint32_t flags, uint32_t tile_mod,
                                              uint32_t tile_rem, uint8_t* out, uint64_t cap, const uint64_t* tile_at);
/* Windowed / reduced-resolution decode, tile-granular: `window` = x0,y0,x1,y1 on the full-resolution
 * canvas (NULL: whole image), `reduce` = highest resolutions to drop.  *cp becomes a VIRTUAL coding: the image made of the
 * tiles the window touches, at 1 / 2^reduce of the resolution -- decode it with b2k_decode(cp, blocks, ..., cs, ...) into
 * planes of (cp->x1 - cp->x0) x (cp->y1 - cp->y0) samples and crop: window column x (reduced resolution, x >=
 * ceil(wx0 / 2^reduce)) is plane column x - cp->x0.  Only the wanted tiles' packets are parsed and decoded.  Return as
 * b2k_codestream_parse (reduce > 0 needs a tile grid aligned to 2^reduce: 1 = not handled otherwise). */
B2K_API int64_t b2k_codestream_parse_window(const uint8_t* cs, uint64_t len, const uint32_t* window, uint32_t reduce, b2k_coding* cp,
                                            b2k_block* blocks, uint64_t cap_blocks);

/* JPH container (JP2 boxes, brand 'jph '): wrap a codestream / find the codestream inside a .jph / .jp2 file
 * (a raw codestream is accepted as it is). */
B2K_API int64_t b2k_jph_wrap(const b2k_coding* cp, const uint8_t* cs, uint64_t cs_len, uint8_t* out, uint64_t cap);
B2K_API int32_t b2k_jph_codestream(const uint8_t* file, uint64_t len, uint64_t* offset, uint64_t* length);

/* ---- streaming: `depth` frames in flight on one GPU, so that frame k+1's host->device copies
 * overlap frame k's kernels and device->host copies.  Each of the `depth` workers is a host thread with its own engine;
 * a submit hands the frame to an idle worker (and blocks while all are busy); results come back through the callback on
 * that worker's thread, possibly out of submission order.  The caller's planes / code-stream bytes are NOT copied: they
 * must stay valid and unchanged until the frame's callback has run.
 *   on_encoded: `result` is valid during the call; return non-zero to keep it (then free it with b2k_result_free).
 * b2k_stream_end drains the stream, joins the workers, frees everything; returns the first device error (< 0) or 0. */
typedef struct b2k_stream b2k_stream;
typedef int32_t (*b2k_encoded_fn)(void* user, void* frame_user, b2k_result* result, int32_t status);
typedef void (*b2k_decoded_fn)(void* user, void* frame_user, int32_t status);
#define B2K_SAMPLES_U16_INTERLEAVED 0x102u /* sample_bytes of an encode stream fed b2k_encode16_interleaved frames: planes[0] = pixels */
B2K_API int32_t b2k_stream_encode_begin(int32_t device, const b2k_coding* cp, uint32_t depth, uint32_t sample_bytes /* 2, 4 or B2K_SAMPLES_U16_INTERLEAVED */,
                                        b2k_encoded_fn on_encoded, void* user, b2k_stream** out);
B2K_API int32_t b2k_stream_encode_submit(b2k_stream* s, const void* const* planes, const uint32_t* strides, void* frame_user);
B2K_API int32_t b2k_stream_decode_begin(int32_t device, uint32_t depth, uint32_t sample_bytes /* 2 or 4 */, b2k_decoded_fn on_decoded,
                                        void* user, b2k_stream** out);
B2K_API int32_t b2k_stream_decode_submit(b2k_stream* s, const b2k_coding* cp, const b2k_block* blocks, uint64_t num_blocks,
                                         const uint8_t* bytes, uint64_t num_bytes, void* const* planes, const uint32_t* strides,
                                         void* frame_user);
B2K_API int32_t b2k_stream_decode_submit_codestream(b2k_stream* s, const uint8_t* codestream, uint64_t length, uint32_t numcomps,
                                                    void* const* planes, const uint32_t* strides, void* frame_user);
B2K_API int32_t b2k_stream_end(b2k_stream* s);

/* launches issued by this library since engine creation (bench.py "gpu_launches") */
B2K_API uint64_t b2k_launch_count(void);
/* per-kernel timing of the last forward()/inverse(): ms of the level-1 kernel and algorithmic
 * bytes it moved (for the roofline line of bench.py) */
B2K_API int32_t b2k_job_last_kernel_stats(const b2k_device_job* j, int which, float* ms, uint64_t* alg_bytes);

#ifdef __cplusplus
}
#endif
#endif /* GROK_B200_H */
