"""ctypes mirrors of the gpup_* structs of include/grok_b200.h (binary-identical to Grok's
plugin/gpup/gpu_plugin_shared.h; sizes checked in tests/test_host.py)."""
import ctypes as C

import grok_b200 as G


class GpupImageComp(C.Structure):
    _fields_ = [("x0", C.c_uint32), ("y0", C.c_uint32), ("w", C.c_uint32), ("stride", C.c_uint32), ("h", C.c_uint32),
                ("dx", C.c_uint8), ("dy", C.c_uint8), ("prec", C.c_uint8), ("sgnd", C.c_bool),
                ("data", C.POINTER(C.c_int32)), ("owns_data", C.c_bool)]


class GpupImage(C.Structure):
    _fields_ = [("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32),
                ("numcomps", C.c_uint16), ("color_space", C.c_int32), ("comps", C.POINTER(GpupImageComp))]


class GpupPass(C.Structure):
    _fields_ = [("distortionDecrease", C.c_double), ("rate", C.c_size_t), ("length", C.c_size_t)]


class GpupCodeBlock(C.Structure):
    _fields_ = [("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32),
                ("contextStream", C.c_void_p), ("numPix", C.c_uint32), ("compressedData", C.POINTER(C.c_uint8)),
                ("compressedDataLength", C.c_uint32), ("numBitPlanes", C.c_uint8), ("numPasses", C.c_size_t),
                ("passes", GpupPass * G.GPUP_MAX_PASSES), ("sortedIndex", C.c_uint)]


class GpupPrecinct(C.Structure):
    _fields_ = [("numBlocks", C.c_uint64), ("blocks", C.POINTER(C.POINTER(GpupCodeBlock)))]


class GpupBand(C.Structure):
    _fields_ = [("orientation", C.c_uint8), ("numPrecincts", C.c_uint64),
                ("precincts", C.POINTER(C.POINTER(GpupPrecinct))), ("stepsize", C.c_float)]


class GpupResolution(C.Structure):
    _fields_ = [("level", C.c_size_t), ("numBands", C.c_size_t), ("band", C.POINTER(C.POINTER(GpupBand)))]


class GpupTileComponent(C.Structure):
    _fields_ = [("numResolutions", C.c_size_t), ("resolutions", C.POINTER(C.POINTER(GpupResolution)))]


class GpupTile(C.Structure):
    _fields_ = [("decompress_flags", C.c_uint32), ("numComponents", C.c_size_t),
                ("tileComponents", C.POINTER(C.POINTER(GpupTileComponent)))]
