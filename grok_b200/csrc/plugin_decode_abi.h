/*
 * grok_b200/csrc/plugin_decode_abi.h -- the one C++-ABI type of Grok's plugin contract.
 *
 * plugin_decompress() hands the host a callback argument that contains std::string members
 * (reference: src/lib/core/plugin/plugin_interface.h L78-115), so it cannot live in the C header
 * include/grok_b200.h.  The struct below restates that layout member for member; host and plugin
 * must be built against the same libstdc++ (they are: Grok dlopens the plugin in-process).
 */
#pragma once
#include <cstring>
#include <string>
#include "../../include/grok_b200.h"

struct PluginDecodeCallbackInfo
{
  PluginDecodeCallbackInfo() : PluginDecodeCallbackInfo("", "", nullptr, 0, 0) {}
  PluginDecodeCallbackInfo(std::string input, std::string output, gpup_decompress_params* decompressorParameters,
                           gpup_codec_fmt format, uint32_t flags)
      : deviceId(0), init_decompressors_func(nullptr), inputFile(input), outputFile(output), decod_format(format),
        cod_format(0), codec(nullptr), decompressor_parameters(decompressorParameters), image(nullptr),
        plugin_owns_image(false), tile(nullptr), error_code(0), decompress_flags(flags), user_data(nullptr),
        format_private(nullptr), codestream(nullptr), codestreamLength(0), frameUser(nullptr)
  {
    memset(&header_info, 0, sizeof(header_info));
  }
  size_t deviceId;
  GPUP_INIT_DECOMPRESSORS init_decompressors_func;
  std::string inputFile;
  std::string outputFile;
  gpup_codec_fmt decod_format;
  gpup_file_fmt cod_format;
  void* codec;
  gpup_decompress_params* decompressor_parameters;
  gpup_header_info header_info;
  gpup_image* image;
  bool plugin_owns_image;
  gpup_tile* tile;
  int32_t error_code;
  uint32_t decompress_flags;
  void* user_data;
  void* format_private;
  const uint8_t* codestream;
  size_t codestreamLength;
  void* frameUser;
};

typedef int32_t (*PLUGIN_DECODE_USER_CALLBACK)(PluginDecodeCallbackInfo* info);

/* plugin_interface.h L117-120 */
extern "C" B2K_API int32_t plugin_decompress(gpup_decompress_params* decoding_parameters, PLUGIN_DECODE_USER_CALLBACK userCallback);
/* plugin_interface.h L130-131 */
extern "C" B2K_API int32_t plugin_batch_decompress_memory_begin(gpup_batch_decompress_memory_info* info,
                                                                PLUGIN_DECODE_USER_CALLBACK userCallback);
