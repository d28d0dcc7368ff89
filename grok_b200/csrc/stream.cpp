/*
 * grok_b200/csrc/stream.cpp -- streaming encode / decode: several frames in flight on one GPU
 * so that frame k+1's host->device copies overlap frame k's kernels and device->host copies.
 *
 * Replaces (reference): the in-memory batch interface of the accelerator plugin
 *   gpup_batch_memory_begin / _submit / _submit_planes / _end      grok.cpp L1538-1545, L1655-1857
 *   (frame shape fixed at begin, frames handed over one by one, results delivered through a callback on plugin
 *   threads, callbacks run concurrently: grok.cpp L1620-1653)
 * and offers the same shape for decoding (cf. plugin_batch_decompress_memory_begin / _end, grok.cpp L2094-2188).
 *
 * Design: a stream owns `depth` workers; a worker = one host thread + one b2k_engine (its own CUDA streams, its own
 * cached job: device buffers and plans of the stream's coding).  A submitted frame goes to the first idle worker,
 * which runs the ordinary chunk-pipelined b2k_encode / b2k_decode on its engine -- engines are independent (engine.cu:
 * per-engine lock and job cache), so the copy engines and SMs see `depth` frames at once: the PCIe legs of
 * neighbouring frames overlap in both directions, which a single synchronous call cannot do.
 * Completion callbacks run on the worker threads, possibly out of submission order.
 */
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/grok_b200.h"

void b2k_set_error(const char* msg);

namespace {

struct Frame
{
  /* encode */
  const void* planes[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t strides[4] = {0, 0, 0, 0};
  /* decode */
  b2k_coding cp{};
  const b2k_block* blocks = nullptr;
  uint64_t num_blocks = 0;
  const uint8_t* bytes = nullptr;
  uint64_t num_bytes = 0;
  bool codestream = false;
  void* out_planes[4] = {nullptr, nullptr, nullptr, nullptr};
  void* frame_user = nullptr;
};

} // namespace

struct b2k_stream
{
  bool decode = false;
  b2k_coding cp{};
  uint32_t sample_bytes = 4;
  b2k_encoded_fn on_encoded = nullptr;
  b2k_decoded_fn on_decoded = nullptr;
  void* user = nullptr;
  std::vector<b2k_engine*> engines;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_room;
  std::deque<Frame> queue;
  uint32_t depth = 0, in_flight = 0;
  bool closing = false;
  int32_t first_error = 0;
};

namespace {

void worker_loop(b2k_stream* S, b2k_engine* eng)
{
  for(;;)
  {
    Frame f;
    {
      std::unique_lock<std::mutex> lk(S->mu);
      S->cv_work.wait(lk, [&] { return S->closing || !S->queue.empty(); });
      if(S->queue.empty())
        return; /* closing and drained */
      f = S->queue.front();
      S->queue.pop_front();
    }
    int32_t rc;
    if(!S->decode)
    {
      b2k_result* R = nullptr;
      if(S->sample_bytes == B2K_SAMPLES_U16_INTERLEAVED)
        rc = b2k_encode16_interleaved(eng, &S->cp, static_cast<const uint16_t*>(f.planes[0]), f.strides[0], 1, 0, &R);
      else if(S->sample_bytes == 2)
        rc = b2k_encode16(eng, &S->cp, reinterpret_cast<const uint16_t* const*>(f.planes), f.strides, 1, 0, &R);
      else
        rc = b2k_encode(eng, &S->cp, reinterpret_cast<const int32_t* const*>(f.planes), f.strides, 1, 0, &R);
      int32_t keep = 0;
      if(S->on_encoded)
        keep = S->on_encoded(S->user, f.frame_user, R, rc);
      if(R && !keep)
        b2k_result_free(R);
    }
    else
    {
      b2k_coding cp = f.cp;
      const b2k_block* blocks = f.blocks;
      uint64_t nblocks = f.num_blocks;
      std::vector<b2k_block> parsed;
      rc = 0;
      if(f.codestream)
      {
        const int64_t n = b2k_codestream_parse(f.bytes, f.num_bytes, &cp, nullptr, 0);
        if(n <= 1)
          rc = n == 1 || n == 0 ? 1 : -1;
        else
        {
          parsed.resize((size_t)n);
          if(b2k_codestream_parse(f.bytes, f.num_bytes, &cp, parsed.data(), (uint64_t)n) != n)
            rc = -1;
          blocks = parsed.data();
          nblocks = (uint64_t)n;
        }
      }
      if(rc == 0)
      {
        if(S->sample_bytes == 2)
          rc = b2k_decode16(eng, &cp, blocks, nblocks, f.bytes, f.num_bytes, reinterpret_cast<uint16_t* const*>(f.out_planes),
                            f.strides, 1, 0, nullptr);
        else
          rc = b2k_decode(eng, &cp, blocks, nblocks, f.bytes, f.num_bytes, reinterpret_cast<int32_t* const*>(f.out_planes),
                          f.strides, 1, 0, nullptr);
      }
      if(S->on_decoded)
        S->on_decoded(S->user, f.frame_user, rc);
    }
    {
      std::lock_guard<std::mutex> lk(S->mu);
      if(rc < 0 && S->first_error == 0)
        S->first_error = rc;
      S->in_flight--;
    }
    S->cv_room.notify_all();
  }
}

int32_t stream_begin(int32_t device, uint32_t depth, b2k_stream* S, b2k_stream** out)
{
  if(depth < 1)
    depth = 1;
  if(depth > 8)
    depth = 8;
  S->depth = depth;
  for(uint32_t i = 0; i < depth; ++i)
  {
    b2k_engine* e = nullptr;
    if(b2k_engine_create(device, &e) != 0)
    {
      for(b2k_engine* x : S->engines)
        b2k_engine_destroy(x);
      delete S;
      return -1;
    }
    S->engines.push_back(e);
  }
  for(uint32_t i = 0; i < depth; ++i)
    S->workers.emplace_back(worker_loop, S, S->engines[i]);
  *out = S;
  return 0;
}

int32_t stream_submit(b2k_stream* S, const Frame& f)
{
  std::unique_lock<std::mutex> lk(S->mu);
  if(S->closing)
    return -1;
  S->cv_room.wait(lk, [&] { return S->in_flight < S->depth; });
  S->in_flight++;
  S->queue.push_back(f);
  lk.unlock();
  S->cv_work.notify_one();
  return 0;
}

} // namespace

extern "C" int32_t b2k_stream_encode_begin(int32_t device, const b2k_coding* cp, uint32_t depth, uint32_t sample_bytes,
                                           b2k_encoded_fn on_encoded, void* user, b2k_stream** out)
{
  if(!cp || !out || (sample_bytes != 2 && sample_bytes != 4 && sample_bytes != B2K_SAMPLES_U16_INTERLEAVED))
    return -1;
  *out = nullptr;
  b2k_stream* S = new b2k_stream();
  S->decode = false;
  S->cp = *cp;
  S->sample_bytes = sample_bytes;
  S->on_encoded = on_encoded;
  S->user = user;
  return stream_begin(device, depth, S, out);
}

extern "C" int32_t b2k_stream_encode_submit(b2k_stream* S, const void* const* planes, const uint32_t* strides, void* frame_user)
{
  if(!S || S->decode || !planes || !strides)
    return -1;
  Frame f;
  const uint16_t nplanes = S->sample_bytes == B2K_SAMPLES_U16_INTERLEAVED ? 1 : S->cp.numcomps;
  for(uint16_t c = 0; c < nplanes && c < 4; ++c)
  {
    f.planes[c] = planes[c];
    f.strides[c] = strides[c];
  }
  f.frame_user = frame_user;
  return stream_submit(S, f);
}

extern "C" int32_t b2k_stream_decode_begin(int32_t device, uint32_t depth, uint32_t sample_bytes, b2k_decoded_fn on_decoded,
                                           void* user, b2k_stream** out)
{
  if(!out || (sample_bytes != 2 && sample_bytes != 4))
    return -1;
  *out = nullptr;
  b2k_stream* S = new b2k_stream();
  S->decode = true;
  S->sample_bytes = sample_bytes;
  S->on_decoded = on_decoded;
  S->user = user;
  return stream_begin(device, depth, S, out);
}

extern "C" int32_t b2k_stream_decode_submit(b2k_stream* S, const b2k_coding* cp, const b2k_block* blocks, uint64_t num_blocks,
                                            const uint8_t* bytes, uint64_t num_bytes, void* const* planes, const uint32_t* strides,
                                            void* frame_user)
{
  if(!S || !S->decode || !cp || !blocks || !planes || !strides)
    return -1;
  Frame f;
  f.cp = *cp;
  f.blocks = blocks;
  f.num_blocks = num_blocks;
  f.bytes = bytes;
  f.num_bytes = num_bytes;
  for(uint16_t c = 0; c < cp->numcomps && c < 4; ++c)
  {
    f.out_planes[c] = planes[c];
    f.strides[c] = strides[c];
  }
  f.frame_user = frame_user;
  return stream_submit(S, f);
}

extern "C" int32_t b2k_stream_decode_submit_codestream(b2k_stream* S, const uint8_t* codestream, uint64_t length,
                                                       uint32_t numcomps, void* const* planes, const uint32_t* strides,
                                                       void* frame_user)
{
  if(!S || !S->decode || !codestream || !planes || !strides || numcomps < 1 || numcomps > 4)
    return -1;
  Frame f;
  f.codestream = true;
  f.bytes = codestream;
  f.num_bytes = length;
  for(uint32_t c = 0; c < numcomps; ++c)
  {
    f.out_planes[c] = planes[c];
    f.strides[c] = strides[c];
  }
  f.frame_user = frame_user;
  return stream_submit(S, f);
}

extern "C" int32_t b2k_stream_end(b2k_stream* S)
{
  if(!S)
    return -1;
  {
    std::lock_guard<std::mutex> lk(S->mu);
    S->closing = true;
  }
  S->cv_work.notify_all();
  for(std::thread& t : S->workers)
    t.join();
  for(b2k_engine* e : S->engines)
    b2k_engine_destroy(e);
  const int32_t rc = S->first_error;
  delete S;
  return rc;
}
