/*
 * oracle/ref_shim/ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY.
 * The inverse-wavelet translation units of the reference (wavelet/WaveletReverse*.cpp) are
 * compiled into oracle/_ref only for their self-contained single-thread benchmark hooks
 * (grk_bench_dwt_53 / grk_bench_dwt_97, WaveletReverse.h L111-118).  The rest of those classes
 * refers to scheduler / logger objects that live in translation units we do not build; the
 * definitions below satisfy the linker and are never reached by the hooks.
 */
#include <cstdlib>
/* the header prerequisites of WaveletReverse.h, in the order wavelet/WaveletReverse.cpp L29-76 pulls them */
#include "hwy_arm_disable_targets.h"
#include "TFSingleton.h"
#include "grk_restrict.h"
#include "simd.h"
#include "CodeStreamLimits.h"
#include "TileWindow.h"
#include "Quantizer.h"
#include "Logger.h"
#include "buffer.h"
#include "GrkObjectWrapper.h"
#include "ISparseCanvas.h"
#include "FlowComponent.h"
#include "IStream.h"
#include "FetchCommon.h"
#include "TPFetchSeq.h"
#include "GrkImageMeta.h"
#include "GrkImage.h"
#include "MarkerParser.h"
#include "PLMarker.h"
#include "SIZMarker.h"
#include "PPMMarker.h"
#include <chrono>
#include <limits>
namespace grk
{
struct ITileProcessor;
}
#include "CodeStream.h"
#include "PacketLengthCache.h"
#include "ICoder.h"
#include "CoderPool.h"
#include "BitIO.h"
#include "ImageComponentFlow.h"
#include "TagTree.h"
#include "CodeblockCompress.h"
#include "CodeblockDecompress.h"
#include "Precinct.h"
#include "Subband.h"
#include "Resolution.h"
#include "TileFutureManager.h"
#include "FlowComponent.h"
#include "CodecScheduler.h"
#include "TileComponentWindow.h"
#include "WaveletReverse.h"
#include "TileComponent.h"
#include "ITileProcessor.h"
#include "DecompressScheduler.h"

namespace grk
{
struct NullLogger : public ILogger
{
  void info(const char*, ...) override {}
  void warn(const char*, ...) override {}
  void error(const char*, ...) override {}
  void debug(const char*, ...) override {}
  void trace(const char*, ...) override {}
};
static NullLogger g_null_logger;
ILogger& grklog = g_null_logger;

ImageComponentFlow* SchedulerStandard::getImageComponentFlow(uint16_t) { abort(); }
Resflow* ImageComponentFlow::getResflow(uint8_t) { abort(); }
bool WaveletReverse::tile_16_97(void) { abort(); }
bool WaveletReverse::decompressPartial() { abort(); }
} // namespace grk
