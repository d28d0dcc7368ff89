/*
 * grok_b200/csrc/plugin_batch.cpp -- the STOCK in-memory batch compress symbols of the accelerator plugin on top of
 * the streaming engine (stream.cpp):
 *   gpup_batch_memory_begin          grok.cpp L1655-1709 (typedefs L1538-1545); gpup_batch_memory_info gpu_plugin_shared.h L452-473
 *   gpup_batch_memory_submit         grok.cpp L1777-1840: one frame, pixel-interleaved little-endian 16-bit samples
 *                                    ("the plugin copies before it returns")
 *   gpup_batch_memory_submit_planes  L1779-1795: GPUP_SOURCE_RGB48LE = the same interleaved layout with a row stride
 *   gpup_batch_memory_end            L1842-1853: drains the batch
 * Every finished frame comes back through the host's GPUP_COMPRESS_USER_CALLBACK (gpu_plugin_shared.h L428-442) with the
 * stock gpup_tile tree, on a plugin thread; the host runs T2 inside it (batchMemoryEncodeCallback, grok.cpp L1620-1653).
 * Frames stay pixel-interleaved across PCIe and are unpacked into planes by a device kernel ("on-device input
 * unpack"); the host only does the copy the contract requires.
 * The stock contract is whole image = one tile; YUV sources (on-device chroma upsampling + matrix) and the on-device
 * X'Y'Z' transform are declined (begin returns 1: the host keeps that batch on the CPU).
 */
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

#include "../../include/grok_b200.h"

void b2k_plugin_log(int level, const char* fmt, ...);
void b2k_host_parallel(size_t n, const std::function<void(size_t)>& fn);
int32_t b2k_plugin_device(void);
void b2k_gpup_tile_free_tree(gpup_tile* tile);

namespace {

struct Slot
{
  uint16_t* pixels = nullptr; /* pinned, pixel-interleaved as submitted (rows packed: stride = numcomps * width) */
  void* host_data = nullptr;
  bool busy = false;
};

struct Batch
{
  bool running = false;
  b2k_stream* stream = nullptr;
  b2k_coding cp{};
  gpup_compress_params params{};
  GPUP_COMPRESS_USER_CALLBACK callback = nullptr;
  uint32_t w = 0, h = 0, nc = 0, prec = 0;
  std::vector<Slot> slots;
  std::mutex mu;
  std::condition_variable cv;
};
Batch g_batch;

int32_t on_encoded(void* user, void* frame_user, b2k_result* R, int32_t status)
{
  Batch* B = static_cast<Batch*>(user);
  Slot* slot = static_cast<Slot*>(frame_user);
  gpup_image_comp comps[4];
  memset(comps, 0, sizeof(comps));
  gpup_image image;
  memset(&image, 0, sizeof(image));
  image.x1 = B->w;
  image.y1 = B->h;
  image.numcomps = (uint16_t)B->nc;
  image.comps = comps;
  for(uint32_t c = 0; c < B->nc; ++c)
  {
    comps[c].w = B->w;
    comps[c].h = B->h;
    comps[c].stride = B->w;
    comps[c].dx = comps[c].dy = 1;
    comps[c].prec = (uint8_t)B->prec;
  }
  gpup_compress_callback_info info;
  memset(&info, 0, sizeof(info));
  info.compressor_parameters = &B->params;
  info.image = &image;
  info.host_data = slot->host_data;
  gpup_tile* tile = nullptr;
  if(status == 0 && R)
    tile = b2k_result_to_gpup_tile(&B->cp, R, 0);
  info.tile = tile;
  info.error_code = (status == 0 && tile) ? 0u : 1u;
  if(B->callback)
    B->callback(&info); /* the host packetises here; the coded bytes stay valid until we return */
  if(tile)
    b2k_gpup_tile_free_tree(tile);
  {
    std::lock_guard<std::mutex> lk(B->mu);
    slot->busy = false;
  }
  B->cv.notify_all();
  return 0; /* the stream frees the result */
}

Slot* take_slot(Batch& B)
{
  std::unique_lock<std::mutex> lk(B.mu);
  Slot* s = nullptr;
  B.cv.wait(lk, [&] {
    for(Slot& x : B.slots)
      if(!x.busy)
      {
        s = &x;
        return true;
      }
    return false;
  });
  s->busy = true;
  return s;
}

} // namespace

extern "C" int32_t gpup_batch_memory_begin(gpup_batch_memory_info* info)
{
  Batch& B = g_batch;
  if(!info || !info->compress_parameters || !info->callback)
    return -1;
  if(B.running)
    return -1;
  if(info->source_format != GPUP_SOURCE_PLANAR_RGB && info->source_format != GPUP_SOURCE_RGB48LE)
    return 1; /* YUV sources: upsampling + matrix stay on the host */
  if(info->source_prec != info->prec || info->prec > 16 || info->numcomps < 1 || info->numcomps > 4)
    return 1;
  /* the coding the stock parameters ask for, image = one tile */
  gpup_image_comp comps[4];
  memset(comps, 0, sizeof(comps));
  gpup_image image;
  memset(&image, 0, sizeof(image));
  static int32_t dummy;
  image.x1 = info->width;
  image.y1 = info->height;
  image.numcomps = (uint16_t)info->numcomps;
  image.comps = comps;
  for(uint32_t c = 0; c < info->numcomps; ++c)
  {
    comps[c].w = info->width;
    comps[c].h = info->height;
    comps[c].stride = info->width;
    comps[c].dx = comps[c].dy = 1;
    comps[c].prec = (uint8_t)info->prec;
    comps[c].data = &dummy;
  }
  if(b2k_coding_from_gpup(info->compress_parameters, &image, 0, &B.cp) != 0)
    return 1;
  B.params = *info->compress_parameters;
  B.callback = info->callback;
  B.w = info->width;
  B.h = info->height;
  B.nc = info->numcomps;
  B.prec = info->prec;
  const uint32_t depth = 3;
  if(b2k_stream_encode_begin(b2k_plugin_device(), &B.cp, depth, B2K_SAMPLES_U16_INTERLEAVED, on_encoded, &B, &B.stream) != 0)
  {
    b2k_plugin_log(2, "batch: no engine: %s", b2k_last_error());
    return -1;
  }
  B.slots.assign(depth + 1, Slot());
  for(Slot& s : B.slots)
  {
    s.pixels = static_cast<uint16_t*>(b2k_host_alloc((size_t)B.w * B.h * B.nc * sizeof(uint16_t)));
    if(!s.pixels)
    {
      b2k_plugin_log(2, "batch: pinned staging allocation failed");
      gpup_batch_memory_end();
      return -1;
    }
  }
  info->xyz_on_device = false;
  B.running = true;
  return 0;
}

/* planes[0] = interleaved 16-bit little-endian samples, row stride in bytes */
extern "C" bool gpup_batch_memory_submit_planes(const uint8_t* const planes[3], const size_t stride_bytes[3], void* host_data)
{
  Batch& B = g_batch;
  if(!B.running || !planes || !planes[0] || !stride_bytes)
    return false;
  Slot* slot = take_slot(B);
  slot->host_data = host_data;
  const uint32_t w = B.w, h = B.h, nc = B.nc;
  const uint8_t* base = planes[0];
  const size_t stride = stride_bytes[0];
  /* the copy the contract asks for ("the frame is copied before this returns") is a plain copy into pinned memory,
     rows in parallel; the samples are split into planes on the device (b2k_encode16_interleaved) */
  const size_t row_bytes = (size_t)w * nc * sizeof(uint16_t);
  const size_t rows_per_task = 64;
  b2k_host_parallel((h + rows_per_task - 1) / rows_per_task, [&](size_t t) {
    const uint32_t y0 = (uint32_t)(t * rows_per_task), y1 = y0 + rows_per_task < h ? (uint32_t)(y0 + rows_per_task) : h;
    uint8_t* dst = reinterpret_cast<uint8_t*>(slot->pixels) + (size_t)y0 * row_bytes;
    if(stride == row_bytes)
      memcpy(dst, base + (size_t)y0 * stride, (size_t)(y1 - y0) * row_bytes);
    else
      for(uint32_t y = y0; y < y1; ++y)
        memcpy(dst + (size_t)(y - y0) * row_bytes, base + (size_t)y * stride, row_bytes);
  });
  const void* p[4] = {slot->pixels, nullptr, nullptr, nullptr};
  const uint32_t strides[4] = {w * nc, 0, 0, 0};
  if(b2k_stream_encode_submit(B.stream, p, strides, slot) != 0)
  {
    std::lock_guard<std::mutex> lk(B.mu);
    slot->busy = false;
    return false;
  }
  return true;
}

extern "C" bool gpup_batch_memory_submit(const uint8_t* packed, void* host_data)
{
  Batch& B = g_batch;
  if(!B.running || !packed)
    return false;
  const uint8_t* planes[3] = {packed, nullptr, nullptr};
  const size_t strides[3] = {(size_t)B.w * B.nc * sizeof(uint16_t), 0, 0};
  return gpup_batch_memory_submit_planes(planes, strides, host_data);
}

extern "C" bool gpup_batch_memory_end(void)
{
  Batch& B = g_batch;
  int32_t rc = 0;
  if(B.stream)
    rc = b2k_stream_end(B.stream); /* drains: every callback has returned */
  B.stream = nullptr;
  for(Slot& s : B.slots)
    if(s.pixels)
    {
      b2k_host_free(s.pixels);
      s.pixels = nullptr;
    }
  B.slots.clear();
  const bool was = B.running;
  B.running = false;
  return was && rc == 0;
}
