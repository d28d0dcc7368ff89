"""grok_b200 -- host-side Python mirror of the B200 JPEG 2000 tile engine's C ABI.

The product is ``libgrokj2k_plugin.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/grok_b200.h``); this module is only the ctypes doorway tests, ``bench.py`` and Python
hosts use.  It mirrors the reference's plugin surface (``src/lib/core/plugin/plugin_interface.h``,
``gpup/gpu_plugin_shared.h``): same entry-point names, argument meaning and return convention
(0 handled, >0 not handled -> host CPU path, <0 device failure).

There is NO CPU fallback here: if the shared library is missing or no CUDA device is present the
calls raise.  (The CPU oracle lives under ``oracle/`` and is test infrastructure only.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2K_LIB") or os.path.join(_HERE, "libgrokj2k_plugin.so")   # B2K_LIB: an experimental build (tools/build_variant.py)

GPUP_MAX_PASSES = 3 * (16 + 7) - 2


class Coding(C.Structure):
    """b2k_coding (include/grok_b200.h)."""
    _fields_ = [("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32),
                ("tx0", C.c_uint32), ("ty0", C.c_uint32), ("tw", C.c_uint32), ("th", C.c_uint32),
                ("numcomps", C.c_uint16), ("prec", C.c_uint8), ("sgnd", C.c_uint8),
                ("numres", C.c_uint8), ("cblkw_exp", C.c_uint8), ("cblkh_exp", C.c_uint8),
                ("irreversible", C.c_uint8), ("mct", C.c_uint8), ("numgbits", C.c_uint8),
                ("prcw_exp", C.c_uint8 * 33), ("prch_exp", C.c_uint8 * 33), ("cblk_sty", C.c_uint8),
                ("qcd_explicit", C.c_uint8), ("qcd_expn", C.c_uint8 * 97), ("qcd_mant", C.c_uint16 * 97)]


class Block(C.Structure):
    """b2k_block."""
    _fields_ = [("tile", C.c_uint32), ("comp", C.c_uint16), ("resno", C.c_uint8), ("band_index", C.c_uint8),
                ("orient", C.c_uint8), ("kmax", C.c_uint8), ("numbps", C.c_uint8), ("numpasses", C.c_uint8),
                ("precno", C.c_uint32), ("cblkno", C.c_uint32),
                ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32),
                ("buf_x", C.c_uint32), ("buf_y", C.c_uint32), ("length", C.c_uint32),
                ("offset", C.c_uint64), ("stepsize", C.c_float), ("length2", C.c_uint32)]


class Result(C.Structure):
    """b2k_result."""
    _fields_ = [("num_blocks", C.c_uint64), ("blocks", C.POINTER(Block)), ("bytes", C.POINTER(C.c_uint8)),
                ("num_bytes", C.c_uint64), ("num_tiles", C.c_uint32),
                ("ms_h2d", C.c_double), ("ms_dwt", C.c_double), ("ms_t1", C.c_double), ("ms_d2h", C.c_double),
                ("ms_total", C.c_double)]


BLOCK_DTYPE = np.dtype([("tile", "<u4"), ("comp", "<u2"), ("resno", "u1"), ("band_index", "u1"), ("orient", "u1"),
                        ("kmax", "u1"), ("numbps", "u1"), ("numpasses", "u1"), ("precno", "<u4"), ("cblkno", "<u4"),
                        ("x0", "<u4"), ("y0", "<u4"), ("x1", "<u4"), ("y1", "<u4"), ("buf_x", "<u4"), ("buf_y", "<u4"),
                        ("length", "<u4"), ("offset", "<u8"), ("stepsize", "<f4"), ("length2", "<u4")])
assert BLOCK_DTYPE.itemsize == C.sizeof(Block), (BLOCK_DTYPE.itemsize, C.sizeof(Block))

# every symbol include/grok_b200.h declares
# plugin_decompress (C++-ABI callback struct) is declared in csrc/plugin_decode_abi.h, not in the C header
EXPORTS = ["minpf_post_load_plugin", "plugin_init", "plugin_get_debug_state", "gpup_encode_mem", "gpup_tile_free",
           "b2k_engine_create", "b2k_engine_destroy", "b2k_last_error", "b2k_host_alloc", "b2k_host_free",
           "b2k_encode", "b2k_encode16", "b2k_encode16_interleaved", "b2k_result_free", "b2k_decode", "b2k_decode16", "b2k_decode_window", "b2k_enumerate",
           "b2k_result_to_gpup_tile", "b2k_job_create", "b2k_job_destroy", "b2k_job_upload", "b2k_job_forward",
           "b2k_job_t1_encode", "b2k_job_t1_decode", "b2k_job_t1_decode_blocks", "b2k_job_inverse", "b2k_job_roundtrip", "b2k_job_roundtrip_n", "b2k_job_roundtrip_pipelined_n", "b2k_job_download",
           "b2k_job_download_coeffs", "b2k_job_upload_coeffs", "b2k_job_fetch_result", "b2k_job_num_blocks",
           "b2k_launch_count", "b2k_job_last_kernel_stats", "b2k_set_host_threads", "b2k_host_pack_last",
           "b2k_codestream_write", "b2k_codestream_parse", "b2k_codestream_parse_window",
           "b2k_codestream_write_tiles", "b2k_codestream_write_tiles_at", "b2k_codestream_write_header", "b2k_jph_wrap", "b2k_jph_codestream", "b2k_result_merge",
           "gpup_encode_mem_tiles", "gpup_tiles_free", "plugin_decompress_codestream", "b2k_coding_from_gpup",
           "b2k_stream_encode_begin", "b2k_stream_encode_submit", "b2k_stream_decode_begin", "b2k_stream_decode_submit",
           "b2k_stream_decode_submit_codestream", "b2k_stream_end",
           "gpup_batch_memory_begin", "gpup_batch_memory_submit", "gpup_batch_memory_submit_planes", "gpup_batch_memory_end",
           "plugin_decompress", "plugin_batch_decompress_memory_begin", "plugin_batch_decompress_memory_end"]

_lib = None


def lib():
    """Load libgrokj2k_plugin.so (built in-tree by __graft_entry__.build / grok_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the engine)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    pp = C.POINTER(C.c_void_p)
    L.b2k_last_error.restype = C.c_char_p
    L.b2k_engine_create.argtypes = [i32, pp]
    L.b2k_engine_destroy.argtypes = [vp]
    L.b2k_host_alloc.argtypes = [C.c_size_t]
    L.b2k_host_alloc.restype = vp
    L.b2k_host_free.argtypes = [vp]
    L.b2k_encode.argtypes = [vp, C.POINTER(Coding), pp, C.POINTER(u32), u32, u32, C.POINTER(C.POINTER(Result))]
    L.b2k_encode16.argtypes = L.b2k_encode.argtypes
    L.b2k_encode16_interleaved.argtypes = [vp, C.POINTER(Coding), vp, u32, u32, u32, C.POINTER(C.POINTER(Result))]
    L.b2k_result_free.argtypes = [C.POINTER(Result)]
    L.b2k_decode.argtypes = [vp, C.POINTER(Coding), vp, u64, vp, u64, pp, C.POINTER(u32), u32, u32,
                             C.POINTER(C.c_double)]
    L.b2k_decode16.argtypes = L.b2k_decode.argtypes
    L.b2k_enumerate.argtypes = [C.POINTER(Coding), u32, u32, vp, u64]
    L.b2k_enumerate.restype = C.c_int64
    L.b2k_result_to_gpup_tile.argtypes = [C.POINTER(Coding), C.POINTER(Result), u32]
    L.b2k_result_to_gpup_tile.restype = vp
    L.gpup_tile_free.argtypes = [vp]
    L.b2k_job_create.argtypes = [vp, C.POINTER(Coding), u32, u32, pp]
    L.b2k_job_destroy.argtypes = [vp]
    for n in ("b2k_job_upload", "b2k_job_download", "b2k_job_download_coeffs", "b2k_job_upload_coeffs"):
        getattr(L, n).argtypes = [vp, pp, C.POINTER(u32)]
    L.b2k_job_forward.argtypes = [vp, C.POINTER(C.c_float)]
    L.b2k_job_inverse.argtypes = [vp, C.POINTER(C.c_float)]
    L.b2k_job_t1_encode.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(u64)]
    L.b2k_job_t1_decode.argtypes = [vp, C.POINTER(C.c_float)]
    L.b2k_job_t1_decode_blocks.argtypes = [vp, vp, u64, vp, u64, C.POINTER(C.c_float)]
    L.b2k_job_roundtrip.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(u64)]
    L.b2k_job_roundtrip_n.argtypes = [vp, u32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(u64)]
    L.b2k_job_roundtrip_pipelined_n.argtypes = [vp, u32, u32, u32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                                C.POINTER(u64)]
    L.b2k_job_fetch_result.argtypes = [vp, C.POINTER(C.POINTER(Result))]
    L.b2k_job_num_blocks.argtypes = [vp]
    L.b2k_job_num_blocks.restype = u64
    L.b2k_launch_count.restype = u64
    L.b2k_set_host_threads.argtypes = [C.c_int32]
    L.b2k_set_host_threads.restype = C.c_int32
    L.b2k_codestream_write.argtypes = [C.POINTER(Coding), C.POINTER(Result), C.c_uint32, vp, u64]
    L.b2k_codestream_write.restype = C.c_int64
    L.b2k_codestream_parse.argtypes = [vp, u64, C.POINTER(Coding), vp, u64]
    L.b2k_codestream_parse.restype = C.c_int64
    L.b2k_result_merge.argtypes = [C.POINTER(Coding), C.POINTER(C.POINTER(Result)), u32, C.POINTER(C.POINTER(Result))]
    L.b2k_jph_wrap.argtypes = [C.POINTER(Coding), vp, u64, vp, u64]
    L.b2k_jph_wrap.restype = C.c_int64
    L.b2k_jph_codestream.argtypes = [vp, u64, C.POINTER(u64), C.POINTER(u64)]
    L.b2k_host_pack_last.argtypes = [C.c_int32]
    L.b2k_host_pack_last.restype = C.c_int32
    L.b2k_job_last_kernel_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(u64)]
    _lib = L
    return L


class EngineError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise EngineError("%s -> %d: %s" % (what, rc, (lib().b2k_last_error() or b"").decode()))


def make_coding(width, height, numcomps=1, prec=8, sgnd=False, numres=6, tile=None, cblk=(64, 64), irreversible=False,
                mct=None, numgbits=1, origin=(0, 0), tile_origin=None, precincts=None):
    """Convenience constructor; defaults follow grk_compress for an HT (.jph) output:
    6 resolutions (CodeStream.h L43), 64x64 blocks (L40), one guard bit (GrkCompress.cpp L849)."""
    cp = Coding()
    cp.x0, cp.y0 = origin
    cp.x1, cp.y1 = origin[0] + width, origin[1] + height
    if tile:
        cp.tw, cp.th = tile
        cp.tx0, cp.ty0 = tile_origin if tile_origin else origin
    cp.numcomps, cp.prec, cp.sgnd, cp.numres = numcomps, prec, int(sgnd), numres
    cp.cblkw_exp, cp.cblkh_exp = int(np.log2(cblk[0])), int(np.log2(cblk[1]))
    cp.irreversible = int(irreversible)
    cp.mct = int(numcomps >= 3) if mct is None else int(mct)
    cp.numgbits = numgbits
    for r in range(33):
        cp.prcw_exp[r] = 15
        cp.prch_exp[r] = 15
    if precincts:  # [(w, h)] per resolution, coarsest first; the last entry repeats
        for r in range(numres):
            pw, ph = precincts[min(r, len(precincts) - 1)]
            cp.prcw_exp[r], cp.prch_exp[r] = int(np.log2(pw)), int(np.log2(ph))
    return cp


def _plane_ptrs(planes):
    n = len(planes)
    arr = (C.c_void_p * n)(*[p.ctypes.data for p in planes])
    strides = (C.c_uint32 * n)(*[p.strides[0] // p.itemsize for p in planes])
    return arr, strides


def set_host_threads(n):
    """Host threads that narrow/widen int32 planes to 16-bit PCIe containers (0 = off, <0 = default)."""
    return int(lib().b2k_set_host_threads(int(n)))


def host_pack_last():
    """(encode, decode): 1 if the last int32 call went through 16-bit host packing, 0 direct, -1 none yet."""
    return int(lib().b2k_host_pack_last(0)), int(lib().b2k_host_pack_last(1))


CS_TLM, CS_PLT, CS_TPARTS_R, CS_SOP, CS_EPH = 1, 2, 4, 16, 32
LRCP, RLCP, RPCL, PCRL, CPRL = range(5)


def CS_PROG(n):
    return (n & 7) << 8


def codestream_write(cp, blocks, data, flags=CS_TLM | CS_PLT, num_tiles=None, out=None):
    """HTJ2K codestream (bytes, numpy uint8) from a coding, a full block table (BLOCK_DTYPE) and its byte arena
    -- an EncodeResult's .blocks / .bytes, or tables built elsewhere (tests build them with the oracle).
    out: optional uint8 buffer to write into (e.g. pinned); a view of the written part is returned."""
    blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    r = Result()
    r.num_blocks = len(blocks)
    r.blocks = C.cast(blocks.ctypes.data, C.POINTER(Block))
    r.bytes = C.cast(data.ctypes.data, C.POINTER(C.c_uint8))
    r.num_bytes = len(data)
    r.num_tiles = int(blocks["tile"].max()) + 1 if num_tiles is None else num_tiles
    if out is not None:
        n = lib().b2k_codestream_write(C.byref(cp), C.byref(r), flags, out.ctypes.data, out.size)
        if 0 <= n <= out.size:
            return out[:n]
    else:
        n = lib().b2k_codestream_write(C.byref(cp), C.byref(r), flags, None, 0)
    if n < 0:
        raise EngineError("b2k_codestream_write: " + (lib().b2k_last_error() or b"").decode())
    out = np.zeros(n, np.uint8)
    n2 = lib().b2k_codestream_write(C.byref(cp), C.byref(r), flags, out.ctypes.data, n)
    assert n2 == n
    return out


def codestream_parse(cs):
    """-> (Coding, block table with offsets into cs).  Raises NotHandled for codestreams outside the path's scope."""
    cs = np.ascontiguousarray(cs, dtype=np.uint8)
    cp = Coding()
    n = lib().b2k_codestream_parse(cs.ctypes.data, len(cs), C.byref(cp), None, 0)
    if n < 0 or n == 1:
        raise (NotHandled if n == 1 else EngineError)("b2k_codestream_parse: " + (lib().b2k_last_error() or b"").decode())
    blocks = np.zeros(n, BLOCK_DTYPE)
    m = lib().b2k_codestream_parse(cs.ctypes.data, len(cs), C.byref(cp), blocks.ctypes.data, n)
    if m != n:
        raise (NotHandled if m == 1 else EngineError)("b2k_codestream_parse: " + (lib().b2k_last_error() or b"").decode())
    return cp, blocks


def result_from_tables(blocks, data, num_tiles):
    """A ctypes Result viewing a numpy block table + byte arena (keep both alive while it is in use)."""
    r = Result()
    r.num_blocks = len(blocks)
    r.blocks = C.cast(blocks.ctypes.data, C.POINTER(Block))
    r.bytes = C.cast(data.ctypes.data, C.POINTER(C.c_uint8))
    r.num_bytes = len(data)
    r.num_tiles = num_tiles
    return r


def codestream_write_tiles(cp, blocks, data, flags=CS_TLM | CS_PLT, tile_mod=1, tile_rem=0, out=None, tile_at=None, sizes_only=False):
    """A shard's tiles as finished tile parts (b2k_codestream_write_tiles) -> (bytes, per-tile lengths of the shard's tiles).
    sizes_only: just the lengths.  tile_at (uint64 offsets, one per tile of the shard): write each tile part at out[tile_at[k]:]
    (b2k_codestream_write_tiles_at) and return (out, None)."""
    blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    r = Result()
    r.num_blocks = len(blocks)
    r.blocks = C.cast(blocks.ctypes.data, C.POINTER(Block))
    r.bytes = C.cast(data.ctypes.data, C.POINTER(C.c_uint8))
    r.num_bytes = len(data)
    L = lib()
    L.b2k_codestream_write_tiles.restype = C.c_int64
    L.b2k_codestream_write_tiles.argtypes = [C.POINTER(Coding), C.POINTER(Result), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    g_nx = -(-(cp.x1 - cp.tx0) // cp.tw) if cp.tw else 1
    g_ny = -(-(cp.y1 - cp.ty0) // cp.th) if cp.th else 1
    nmine = len(range(tile_rem, g_nx * g_ny, tile_mod))
    lens = np.zeros(nmine, np.uint64)
    if tile_at is not None:
        L.b2k_codestream_write_tiles_at.restype = C.c_int64
        L.b2k_codestream_write_tiles_at.argtypes = [C.POINTER(Coding), C.POINTER(Result), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        ta = np.ascontiguousarray(tile_at, dtype=np.uint64)
        n = L.b2k_codestream_write_tiles_at(C.byref(cp), C.byref(r), flags, tile_mod, tile_rem, out.ctypes.data, out.size, ta.ctypes.data)
        if n < 0:
            raise EngineError("b2k_codestream_write_tiles_at: " + (L.b2k_last_error() or b"").decode())
        return out, None
    n = L.b2k_codestream_write_tiles(C.byref(cp), C.byref(r), flags, tile_mod, tile_rem, None, 0, lens.ctypes.data)
    if n < 0:
        raise EngineError("b2k_codestream_write_tiles: " + (L.b2k_last_error() or b"").decode())
    if sizes_only:
        return None, lens
    if out is None or out.size < n:
        out = np.zeros(max(n, 1), np.uint8)
    assert L.b2k_codestream_write_tiles(C.byref(cp), C.byref(r), flags, tile_mod, tile_rem, out.ctypes.data, out.size, lens.ctypes.data) == n
    return out[:n], lens


def codestream_write_header(cp, flags, tile_bytes_all):
    """Main header (+ TLM) from the tile-part length of every tile (b2k_codestream_write_header)."""
    tb = np.ascontiguousarray(tile_bytes_all, dtype=np.uint64)
    L = lib()
    L.b2k_codestream_write_header.restype = C.c_int64
    L.b2k_codestream_write_header.argtypes = [C.POINTER(Coding), C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
    n = L.b2k_codestream_write_header(C.byref(cp), flags, tb.ctypes.data, len(tb), None, 0)
    if n < 0:
        raise EngineError("b2k_codestream_write_header: " + (L.b2k_last_error() or b"").decode())
    out = np.zeros(n, np.uint8)
    assert L.b2k_codestream_write_header(C.byref(cp), flags, tb.ctypes.data, len(tb), out.ctypes.data, n) == n
    return out


def merge_shards(cp, shards):
    """shards: [(block table, byte arena)] of ranks 0..n-1 (rank r coded the tiles t % n == r) -> EncodeResult holding
    every tile in enumeration order (b2k_result_merge)."""
    keep = [(np.ascontiguousarray(b, dtype=BLOCK_DTYPE), np.ascontiguousarray(d, dtype=np.uint8)) for b, d in shards]
    rs = [result_from_tables(b, d, 0) for b, d in keep]
    arr = (C.POINTER(Result) * len(rs))(*[C.pointer(r) for r in rs])
    out = C.POINTER(Result)()
    _check(lib().b2k_result_merge(C.byref(cp), arr, len(rs), C.byref(out)), "b2k_result_merge")
    return EncodeResult(out)


def jph_wrap(cp, cs):
    """codestream -> .jph file bytes (JP2 boxes, brand 'jph ')."""
    cs = np.ascontiguousarray(cs, dtype=np.uint8)
    n = lib().b2k_jph_wrap(C.byref(cp), cs.ctypes.data, len(cs), None, 0)
    if n < 0:
        raise EngineError("b2k_jph_wrap: " + (lib().b2k_last_error() or b"").decode())
    out = np.zeros(n, np.uint8)
    assert lib().b2k_jph_wrap(C.byref(cp), cs.ctypes.data, len(cs), out.ctypes.data, n) == n
    return out


def jph_codestream(data):
    """.jph / .jp2 file bytes (or a raw codestream) -> view of the contiguous codestream."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off, n = C.c_uint64(), C.c_uint64()
    if lib().b2k_jph_codestream(data.ctypes.data, len(data), C.byref(off), C.byref(n)) != 0:
        raise EngineError("b2k_jph_codestream: " + (lib().b2k_last_error() or b"").decode())
    return data[off.value:off.value + n.value]


class NotHandled(EngineError):
    pass


def pinned_empty(shape, dtype):
    """numpy array in cudaHostAlloc'ed (pinned) memory; keep the returned array alive, free with
    ``lib().b2k_host_free(arr.ctypes.data)`` (or let the process end)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib().b2k_host_alloc(n)
    if not p:
        raise EngineError("b2k_host_alloc(%d) failed" % n)
    buf = (C.c_uint8 * n).from_address(p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class EncodeResult:
    """Owns a b2k_result; exposes numpy views of the block table and the byte arena."""

    def __init__(self, ptr):
        self._ptr = ptr
        r = ptr.contents
        self.num_blocks = int(r.num_blocks)
        self.num_bytes = int(r.num_bytes)
        self.num_tiles = int(r.num_tiles)
        self.timings = dict(h2d=r.ms_h2d, dwt=r.ms_dwt, t1=r.ms_t1, d2h=r.ms_d2h, total=r.ms_total)
        self.blocks = np.frombuffer((C.c_uint8 * (self.num_blocks * C.sizeof(Block))).from_address(
            C.addressof(r.blocks.contents)), dtype=BLOCK_DTYPE) if self.num_blocks else np.zeros(0, BLOCK_DTYPE)
        self.bytes = np.frombuffer((C.c_uint8 * max(1, self.num_bytes)).from_address(
            C.addressof(r.bytes.contents)), dtype=np.uint8)[:self.num_bytes]

    def block_bytes(self, i):
        b = self.blocks[i]
        return self.bytes[int(b["offset"]):int(b["offset"]) + int(b["length"])]

    def free(self):
        if self._ptr is not None:
            lib().b2k_result_free(self._ptr)
            self._ptr = None
            self.blocks = self.bytes = None

    def __del__(self):
        self.free()


class Engine:
    """b2k_engine: one per process per GPU (one process per GPU is the deployment model)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().b2k_engine_create(device, C.byref(self._h)), "b2k_engine_create")

    def close(self):
        if self._h:
            lib().b2k_engine_destroy(self._h)
            self._h = C.c_void_p()

    def encode(self, cp, planes, tile_mod=1, tile_rem=0):
        """planes: list of 2-D int32 arrays (row stride may exceed width), or uint16 / int16 arrays
        (16-bit containers, b2k_encode16)."""
        ptrs, strides = _plane_ptrs(planes)
        out = C.POINTER(Result)()
        fn = "b2k_encode16" if planes[0].itemsize == 2 else "b2k_encode"
        _check(getattr(lib(), fn)(self._h, C.byref(cp), ptrs, strides, tile_mod, tile_rem, C.byref(out)), fn)
        return EncodeResult(out)

    def encode_interleaved(self, cp, pixels, tile_mod=1, tile_rem=0):
        """pixels: one (H, W, numcomps) uint16 / int16 array (RGB48LE rows; the row stride may exceed W * numcomps
        samples): b2k_encode16_interleaved -- the rows cross PCIe as they are, planes are made on the device."""
        assert pixels.ndim == 3 and pixels.itemsize == 2 and pixels.strides[2] == 2 and pixels.strides[1] == 2 * pixels.shape[2]
        out = C.POINTER(Result)()
        _check(lib().b2k_encode16_interleaved(self._h, C.byref(cp), pixels.ctypes.data, pixels.strides[0] // 2, tile_mod, tile_rem,
                                              C.byref(out)), "b2k_encode16_interleaved")
        return EncodeResult(out)

    def encode_codestream(self, cp, planes, flags=CS_TLM | CS_PLT, out=None):
        """planes -> a complete HTJ2K codestream (numpy uint8): b2k_encode + b2k_codestream_write."""
        res = self.encode(cp, planes)
        try:
            return codestream_write(cp, res.blocks, res.bytes, flags, num_tiles=res.num_tiles, out=out)
        finally:
            res.free()

    def decode_codestream(self, cs, dtype=np.int32, out=None):
        """HTJ2K codestream -> (Coding, list of planes): b2k_codestream_parse + b2k_decode, block bytes read in place."""
        cs = np.ascontiguousarray(cs, dtype=np.uint8)
        cp, blocks = codestream_parse(cs)
        if out is None:
            out = [np.zeros((cp.y1 - cp.y0, cp.x1 - cp.x0), dtype) for _ in range(cp.numcomps)]
        self.decode(cp, blocks, cs, out)
        return cp, out

    def decode_window(self, cs, window=None, reduce=0, dtype=np.int32):
        """Tile-granular windowed / reduced-resolution decode of an HTJ2K codestream (b2k_codestream_parse_window +
        b2k_decode_window: the touched tiles are decoded, the window's pixels alone are copied back).  window = (x0, y0, x1, y1) on the full-resolution canvas or None; returns (virtual Coding,
        planes of the window at 1 / 2**reduce resolution).  The planes are views of pinned buffers the engine keeps and
        reuses for later windows of the same tile-box shape: copy them to keep them."""
        cs = np.ascontiguousarray(cs, dtype=np.uint8)
        L = lib()
        L.b2k_codestream_parse_window.restype = C.c_int64
        L.b2k_codestream_parse_window.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(Coding), C.c_void_p, C.c_uint64]
        win = (C.c_uint32 * 4)(*window) if window is not None else None
        cp = Coding()
        n = L.b2k_codestream_parse_window(cs.ctypes.data, len(cs), win, reduce, C.byref(cp), None, 0)
        if n <= 1:
            raise EngineError("b2k_codestream_parse_window: %d %s" % (n, (L.b2k_last_error() or b"").decode()))
        blocks = np.zeros(n, BLOCK_DTYPE)
        m = L.b2k_codestream_parse_window(cs.ctypes.data, len(cs), win, reduce, C.byref(cp), blocks.ctypes.data, n)
        if m != n:
            raise EngineError("b2k_codestream_parse_window: %d %s" % (m, (L.b2k_last_error() or b"").decode()))
        if window is None:
            rect = (cp.x0, cp.y0, cp.x1, cp.y1)
        else:
            sh = (1 << reduce) - 1
            x0, y0, x1, y1 = [(v + sh) >> reduce for v in window]
            rect = (max(x0, cp.x0), max(y0, cp.y0), min(x1, cp.x1), min(y1, cp.y1))
        # pinned landing planes of the WINDOW's size, kept per shape: only the window's pixels come back over PCIe
        cache = self.__dict__.setdefault("_win_planes", {})
        key = (rect[3] - rect[1], rect[2] - rect[0], cp.numcomps, np.dtype(dtype).str)
        out = cache.get(key)
        if out is None:
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            out = cache[key] = [pinned_empty((key[0], key[1]), dtype) for _ in range(cp.numcomps)]
        ptrs, strides = _plane_ptrs(out)
        ms = C.c_double()
        L.b2k_decode_window.argtypes = [C.c_void_p, C.POINTER(Coding), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32,
                                        C.POINTER(C.c_double)]
        _check(L.b2k_decode_window(self._h, C.byref(cp), blocks.ctypes.data, n, cs.ctypes.data, len(cs), ptrs, strides,
                                   (C.c_uint32 * 4)(*rect), np.dtype(dtype).itemsize, C.byref(ms)), "b2k_decode_window")
        return cp, out

    def decode(self, cp, blocks, data, out_planes, tile_mod=1, tile_rem=0):
        ptrs, strides = _plane_ptrs(out_planes)
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        ms = C.c_double()
        fn = "b2k_decode16" if out_planes[0].itemsize == 2 else "b2k_decode"
        _check(getattr(lib(), fn)(self._h, C.byref(cp), blocks.ctypes.data, len(blocks), data.ctypes.data, len(data),
                                  ptrs, strides, tile_mod, tile_rem, C.byref(ms)), fn)
        return ms.value

    def job(self, cp, tile_mod=1, tile_rem=0):
        return Job(self, cp, tile_mod, tile_rem)


class Job:
    """b2k_device_job: device-resident buffers + per-stage entry points (parity tests, bench `value`)."""

    def __init__(self, eng, cp, tile_mod=1, tile_rem=0):
        self._h = C.c_void_p()
        self.cp = cp
        _check(lib().b2k_job_create(eng._h, C.byref(cp), tile_mod, tile_rem, C.byref(self._h)), "b2k_job_create")

    def close(self):
        if self._h:
            lib().b2k_job_destroy(self._h)
            self._h = C.c_void_p()

    def _planes(self, fn, planes):
        ptrs, strides = _plane_ptrs(planes)
        _check(getattr(lib(), fn)(self._h, ptrs, strides), fn)

    def upload(self, planes):
        self._planes("b2k_job_upload", planes)

    def download(self, planes):
        self._planes("b2k_job_download", planes)

    def download_coeffs(self, planes):
        self._planes("b2k_job_download_coeffs", planes)

    def upload_coeffs(self, planes):
        self._planes("b2k_job_upload_coeffs", planes)

    def forward(self):
        ms = C.c_float()
        _check(lib().b2k_job_forward(self._h, C.byref(ms)), "b2k_job_forward")
        return ms.value

    def inverse(self):
        ms = C.c_float()
        _check(lib().b2k_job_inverse(self._h, C.byref(ms)), "b2k_job_inverse")
        return ms.value

    def t1_encode(self):
        ms, total = C.c_float(), C.c_uint64()
        _check(lib().b2k_job_t1_encode(self._h, C.byref(ms), C.byref(total)), "b2k_job_t1_encode")
        return ms.value, total.value

    def t1_decode(self):
        ms = C.c_float()
        _check(lib().b2k_job_t1_decode(self._h, C.byref(ms)), "b2k_job_t1_decode")
        return ms.value

    def t1_decode_blocks(self, blocks, data):
        """Block-decode a caller-supplied block table (BLOCK_DTYPE) + byte arena into the coefficient planes."""
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        ms = C.c_float()
        _check(lib().b2k_job_t1_decode_blocks(self._h, blocks.ctypes.data, len(blocks), data.ctypes.data, len(data), C.byref(ms)),
               "b2k_job_t1_decode_blocks")
        return ms.value

    def roundtrip_n(self, steps):
        """`steps` round trips queued back to back, one synchronisation.  Returns (total ms, [fwd, enc, dec, inv] ms
        summed over the steps, level-1 DWT kernel ms summed, coded bytes)."""
        ms, st, l1, nb = C.c_float(), (C.c_float * 4)(), C.c_float(), C.c_uint64()
        rc = lib().b2k_job_roundtrip_n(self._h, steps, C.byref(ms), st, C.byref(l1), C.byref(nb))
        if rc == 2:
            rc = lib().b2k_job_roundtrip_n(self._h, steps, C.byref(ms), st, C.byref(l1), C.byref(nb))
        _check(rc, "b2k_job_roundtrip_n")
        return ms.value, [float(v) for v in st], l1.value, int(nb.value)

    def roundtrip_pipelined_n(self, steps, chunks=0, streams=0):
        """`steps` round trips with the block-coder stage pipelined over block ranges on side streams
        (b2k_job_roundtrip_pipelined_n).  Returns (total ms, [fwd, block coder, inv] ms summed, level-1 kernel ms summed, bytes)."""
        ms, st, l1, nb = C.c_float(), (C.c_float * 3)(), C.c_float(), C.c_uint64()
        args = (self._h, steps, chunks, streams, C.byref(ms), st, C.byref(l1), C.byref(nb))
        rc = lib().b2k_job_roundtrip_pipelined_n(*args)
        if rc == 2:
            rc = lib().b2k_job_roundtrip_pipelined_n(*args)
        _check(rc, "b2k_job_roundtrip_pipelined_n")
        return ms.value, [float(v) for v in st], l1.value, int(nb.value)

    def roundtrip(self):
        """forward -> block encode -> block decode -> inverse, device-resident, one synchronisation.
        Returns (total ms, [fwd, enc, dec, inv] ms, coded bytes)."""
        ms, st, nb = C.c_float(), (C.c_float * 4)(), C.c_uint64()
        rc = lib().b2k_job_roundtrip(self._h, C.byref(ms), st, C.byref(nb))
        if rc == 2:
            rc = lib().b2k_job_roundtrip(self._h, C.byref(ms), st, C.byref(nb))
        _check(rc, "b2k_job_roundtrip")
        return ms.value, list(st), nb.value

    def fetch_result(self):
        out = C.POINTER(Result)()
        _check(lib().b2k_job_fetch_result(self._h, C.byref(out)), "b2k_job_fetch_result")
        return EncodeResult(out)

    def num_blocks(self):
        return int(lib().b2k_job_num_blocks(self._h))

    def kernel_stats(self, which=0):
        ms, nb = C.c_float(), C.c_uint64()
        lib().b2k_job_last_kernel_stats(self._h, which, C.byref(ms), C.byref(nb))
        return ms.value, nb.value


def enumerate_blocks(cp, tile_mod=1, tile_rem=0):
    n = lib().b2k_enumerate(C.byref(cp), tile_mod, tile_rem, None, 0)
    if n < 0:
        raise EngineError("b2k_enumerate: " + (lib().b2k_last_error() or b"").decode())
    out = np.zeros(n, BLOCK_DTYPE)
    lib().b2k_enumerate(C.byref(cp), tile_mod, tile_rem, out.ctypes.data, n)
    return out


# ------------------------------------------------------------------------------------------------
# streaming (include/grok_b200.h "streaming", csrc/stream.cpp): `depth` frames in flight on one GPU
# ------------------------------------------------------------------------------------------------
_ENCODED_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(Result), C.c_int32)
_DECODED_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int32)


def _bind_stream():
    L = lib()
    if getattr(L, "_b2k_stream_bound", False):
        return L
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    pp = C.POINTER(C.c_void_p)
    L.b2k_stream_encode_begin.argtypes = [i32, C.POINTER(Coding), u32, u32, _ENCODED_FN, vp, pp]
    L.b2k_stream_encode_submit.argtypes = [vp, pp, C.POINTER(u32), vp]
    L.b2k_stream_decode_begin.argtypes = [i32, u32, u32, _DECODED_FN, vp, pp]
    L.b2k_stream_decode_submit.argtypes = [vp, C.POINTER(Coding), vp, u64, vp, u64, pp, C.POINTER(u32), vp]
    L.b2k_stream_decode_submit_codestream.argtypes = [vp, vp, u64, u32, pp, C.POINTER(u32), vp]
    L.b2k_stream_end.argtypes = [vp]
    L._b2k_stream_bound = True
    return L


class EncodeStream:
    """b2k_stream_encode_*: submit(planes, tag) hands a frame to an idle worker (blocks while `depth` frames are in
    flight); on_encoded(tag, EncodeResult or None, status) runs on a worker thread -- the EncodeResult is the caller's
    (free it).  The planes must stay alive and unchanged until the callback for their frame has run."""

    def __init__(self, cp, depth=3, sample_bytes=4, on_encoded=None, device=0):
        L = _bind_stream()
        self._cp = cp
        self._tags = {}
        self._next = 1
        self._user_cb = on_encoded

        def cb(_user, frame_user, result, status):
            tag, keep = self._tags.pop(int(frame_user or 0), (None, None))
            res = EncodeResult(result) if (status == 0 and result) else None
            if self._user_cb:
                self._user_cb(tag, res, status)
            elif res is not None:
                res.free()
            return 1 if res is not None else 0      # the EncodeResult owns the b2k_result now

        self._cb = _ENCODED_FN(cb)
        self._h = C.c_void_p()
        _check(L.b2k_stream_encode_begin(device, C.byref(cp), depth, sample_bytes, self._cb, None, C.byref(self._h)),
               "b2k_stream_encode_begin")

    def submit(self, planes, tag=None):
        ptrs, strides = _plane_ptrs(planes)
        key = self._next
        self._next += 1
        self._tags[key] = (tag, planes)              # keeps the planes alive until the callback
        _check(lib().b2k_stream_encode_submit(self._h, ptrs, strides, C.c_void_p(key)), "b2k_stream_encode_submit")

    def end(self):
        if self._h:
            rc = lib().b2k_stream_end(self._h)
            self._h = C.c_void_p()
            return rc
        return 0


class DecodeStream:
    """b2k_stream_decode_*: submit(cp, blocks, data, out_planes, tag) / submit_codestream(cs, out_planes, tag);
    on_decoded(tag, status) runs on a worker thread once out_planes hold the pixels."""

    def __init__(self, depth=3, sample_bytes=4, on_decoded=None, device=0):
        L = _bind_stream()
        self._tags = {}
        self._next = 1
        self._user_cb = on_decoded

        def cb(_user, frame_user, status):
            tag = self._tags.pop(int(frame_user or 0), (None,))[0]
            if self._user_cb:
                self._user_cb(tag, status)

        self._cb = _DECODED_FN(cb)
        self._h = C.c_void_p()
        _check(L.b2k_stream_decode_begin(device, depth, sample_bytes, self._cb, None, C.byref(self._h)), "b2k_stream_decode_begin")

    def submit(self, cp, blocks, data, out_planes, tag=None):
        ptrs, strides = _plane_ptrs(out_planes)
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        key = self._next
        self._next += 1
        self._tags[key] = (tag, blocks, data, out_planes, cp)
        _check(lib().b2k_stream_decode_submit(self._h, C.byref(cp), blocks.ctypes.data, len(blocks), data.ctypes.data, len(data),
                                              ptrs, strides, C.c_void_p(key)), "b2k_stream_decode_submit")

    def submit_codestream(self, cs, out_planes, tag=None):
        ptrs, strides = _plane_ptrs(out_planes)
        cs = np.ascontiguousarray(cs, dtype=np.uint8)
        key = self._next
        self._next += 1
        self._tags[key] = (tag, cs, out_planes)
        _check(lib().b2k_stream_decode_submit_codestream(self._h, cs.ctypes.data, len(cs), len(out_planes), ptrs, strides,
                                                         C.c_void_p(key)), "b2k_stream_decode_submit_codestream")

    def end(self):
        if self._h:
            rc = lib().b2k_stream_end(self._h)
            self._h = C.c_void_p()
            return rc
        return 0
