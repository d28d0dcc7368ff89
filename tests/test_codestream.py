"""CPU tests (-m "not gpu") of the codestream writer / parser (grok_b200/csrc/codestream.cpp).
The check that matters is independent: codestreams assembled here -- from blocks the oracle coded, so no GPU is
involved -- are decoded by OpenJPEG (two separate builds: Pillow's and OpenCV's), a decoder that shares nothing
with this repo or with the reference, and must give back the source image exactly (reversible path) or within
the reference's lossy tolerances (irreversible path).  That pins, end to end: marker segments, packet headers
(tag trees, Lblock, HT segment lengths), TLM/PLT, and with them the oracle's DWT / RCT / quantiser / HT coder."""
import io
import os

import numpy as np
import pytest

import grok_b200 as G
import oracle_pipeline as P

PIL_Image = pytest.importorskip("PIL.Image")


def oracle_encode(cp, planes):
    coefs = P.forward(cp, planes)
    table = G.enumerate_blocks(cp)
    blks = P.enumerate_all(cp)
    rects = P.tile_rects(cp)
    assert len(table) == len(blks)
    chunks, off = [], 0
    for i, (t, c, b) in enumerate(blks):
        data = P.encode_block(cp, coefs, rects[t], c, b)
        table[i]["length"], table[i]["offset"], table[i]["numbps"], table[i]["numpasses"] = len(data), off, 1, 1
        chunks.append(data)
        off += len(data)
    return table, np.concatenate(chunks), coefs


def oracle_decode(cp, blocks, cs):
    """parsed block table + codestream -> pixels, all on the oracle"""
    blks = P.enumerate_all(cp)
    rects = P.tile_rects(cp)
    w, h = cp.x1 - cp.x0, cp.y1 - cp.y0
    coefs = [np.zeros((h, w), np.int32) for _ in range(cp.numcomps)]
    for i, (t, c, b) in enumerate(blks):
        bw, bh = b.x1 - b.x0, b.y1 - b.y0
        if bw == 0 or bh == 0 or blocks[i]["length"] == 0:
            continue
        o, n = int(blocks[i]["offset"]), int(blocks[i]["length"])
        win = P.decode_block(cp, cs[o:o + n], c, b, numbps=int(blocks[i]["numbps"]))
        x0, y0 = rects[t][0] - cp.x0, rects[t][1] - cp.y0
        coefs[c][y0 + b.buf_y:y0 + b.buf_y + bh, x0 + b.buf_x:x0 + b.buf_x + bw] = win
    return P.inverse(cp, coefs)


def openjpeg_pillow(cs):
    im = PIL_Image.open(io.BytesIO(cs.tobytes()))
    im.load()
    return np.asarray(im)


def openjpeg_cv2(cs):
    cv2 = pytest.importorskip("cv2")
    a = cv2.imdecode(np.frombuffer(cs.tobytes(), np.uint8), cv2.IMREAD_UNCHANGED)
    assert a is not None, "OpenCV's OpenJPEG could not decode the codestream"
    return a if a.ndim == 2 else a[:, :, ::-1]  # BGR -> RGB


CASES = [
    dict(width=512, height=512, numcomps=1, prec=8),                                              # BASELINE config 1
    dict(width=300, height=200, numcomps=3, prec=8, numres=5, tile=(128, 128)),                   # tiles + RCT
    dict(width=333, height=217, numcomps=3, prec=8, numres=4, origin=(3, 5)),                     # odd canvas origin
    dict(width=260, height=190, numcomps=3, prec=12, numres=5, tile=(100, 64), precincts=[(64, 64), (128, 128)]),
    dict(width=200, height=120, numcomps=1, prec=16, numres=3, cblk=(32, 32)),                    # 16-bit grey
    dict(width=61, height=9, numcomps=1, prec=8, numres=6),                                       # levels run out of samples
    dict(width=130, height=70, numcomps=3, prec=8, numres=3, tile=(64, 64), origin=(64, 33), tile_origin=(1, 1), cblk=(16, 64)),   # tile grid anchored before the image
    dict(width=150, height=110, numcomps=3, prec=8, numres=1, tile=(64, 64)),                     # no wavelet level at all
]


@pytest.mark.parametrize("args", CASES)
@pytest.mark.parametrize("flags", [0, G.CS_TLM | G.CS_PLT])
def test_reversible_codestream_is_decoded_exactly_by_openjpeg(args, flags):
    cp = G.make_coding(**args)
    planes = P.synthetic_image(args["width"], args["height"], args["numcomps"], args["prec"], seed=21,
                               origin=args.get("origin", (0, 0)))
    table, data, _ = oracle_encode(cp, planes)
    cs = G.codestream_write(cp, table, data, flags)
    assert cs[0] == 0xFF and cs[1] == 0x4F and cs[-2] == 0xFF and cs[-1] == 0xD9
    src = planes[0] if len(planes) == 1 else np.stack(planes, axis=-1)
    checked = 0
    if args.get("origin", (0, 0)) == (0, 0):             # OpenCV's wrapper refuses a canvas offset
        got = openjpeg_cv2(cs)
        assert got.shape == src.shape and np.array_equal(got.astype(np.int64), src)
        checked += 1
    if args["prec"] == 8 or args["numcomps"] == 1:       # Pillow keeps more than 8 bits only for grey images
        got = openjpeg_pillow(cs)
        assert got.shape == src.shape and np.array_equal(got.astype(np.int64), src)
        checked += 1
    assert checked
    # and our own parser reads back exactly what went in: coding, block table, byte ranges
    cp2, blocks = G.codestream_parse(cs)
    for f, _ in G.Coding._fields_:
        if f not in ("tx0", "ty0", "tw", "th", "prcw_exp", "prch_exp", "qcd_expn", "qcd_mant"):
            assert getattr(cp2, f) == getattr(cp, f), f
    assert P.tile_rects(cp2) == P.tile_rects(cp)         # an untiled coding comes back as one explicit tile
    assert list(cp2.prcw_exp)[:cp.numres] == list(cp.prcw_exp)[:cp.numres]
    assert list(cp2.prch_exp)[:cp.numres] == list(cp.prch_exp)[:cp.numres]
    for f in ("tile", "comp", "resno", "band_index", "orient", "kmax", "numbps", "numpasses", "precno", "cblkno", "x0", "y0",
              "x1", "y1", "buf_x", "buf_y", "length", "length2"):
        assert np.array_equal(blocks[f], table[f]), f
    for i in np.flatnonzero(table["length"])[::7]:
        o, n = int(blocks[i]["offset"]), int(blocks[i]["length"])
        assert np.array_equal(cs[o:o + n], data[int(table[i]["offset"]):int(table[i]["offset"]) + n])
    rec = oracle_decode(cp2, blocks, cs)
    for a, b in zip(rec, planes):
        assert np.array_equal(a, b)


def test_irreversible_codestream_is_decoded_by_openjpeg_within_tolerance():
    """9/7 + ICT + HT quantiser step sizes as QCD signals them: OpenJPEG's reconstruction agrees with the oracle's
    own decode of the same codestream to within one code, and both are close to the source
    (GrkPluginMemoryTest.cpp L39-52: lossy 8-bit <= 2 codes... here measured against the 12-bit rule: <= 16 codes, PSNR > 50 dB)."""
    w, h = 320, 200
    cp = G.make_coding(w, h, 3, 12, numres=5, irreversible=True)
    planes = P.synthetic_image(w, h, 3, 12, seed=22)
    table, data, _ = oracle_encode(cp, planes)
    cs = G.codestream_write(cp, table, data)
    got = openjpeg_cv2(cs).astype(np.int64)
    cp2, blocks = G.codestream_parse(cs)
    ours = np.stack(oracle_decode(cp2, blocks, cs), axis=-1).astype(np.int64)
    src = np.stack(planes, axis=-1).astype(np.int64)
    assert np.abs(got - ours).max() <= 1
    for rec in (got, ours):
        err = (rec - src).astype(np.float64)
        assert np.abs(err).max() <= 16
        assert 10 * np.log10(4095.0 ** 2 / (err ** 2).mean()) > 50.0


def test_refinement_passes_survive_the_packet_headers():
    """2- and 3-pass HT blocks (cleanup | refinement segments, T.814 B.10.7): lengths of both segments and the pass
    count come back from the parser; OpenJPEG decodes the stream too (it implements SigProp / MagRef)."""
    import oracle_lib as O
    w, h = 96, 80
    cp = G.make_coding(w, h, 1, 8, numres=3, cblk=(32, 32))
    planes = P.synthetic_image(w, h, 1, 8, seed=23)
    coefs = P.forward(cp, planes)
    table = G.enumerate_blocks(cp)
    blks = P.enumerate_all(cp)
    chunks, off = [], 0
    want = np.zeros((h, w), np.int32)
    L = O.lib()
    for i, (t, c, b) in enumerate(blks):
        bw, bh = b.x1 - b.x0, b.y1 - b.y0
        kmax, _, _ = P.band_params(cp, b.resno, b.orient)
        win = np.ascontiguousarray(coefs[c][b.buf_y:b.buf_y + bh, b.buf_x:b.buf_x + bw])
        sm = np.zeros(bw * bh, np.uint32)
        L.orc_ht_pre_rev(win, bw, bw, bh, kmax, sm)
        W = (((sm & 0x7FFFFFFF) << 1) | (sm & 0x80000000)).astype(np.uint32).reshape(bh, bw)
        npass = 1 + i % 3
        s = 1 if (npass > 1 and kmax >= 3) else 0
        npass = npass if s else 1
        mm = kmax - 1 - s
        cup = O.ht_encode(W, mm)
        seg = O.ht_encode_refine(W, mm, npass) if npass > 1 else np.zeros(0, np.uint8)
        rc, dec = O.ht_decode_passes(np.concatenate([cup, seg]), len(seg), npass, mm, bw, bh)
        assert rc == 0
        out = np.zeros((bh, bw), np.int32)
        L.orc_ht_post_rev(dec, bw, bw, bh, kmax, out, bw)
        want[b.buf_y:b.buf_y + bh, b.buf_x:b.buf_x + bw] = out
        table[i]["length"], table[i]["length2"], table[i]["offset"] = len(cup), len(seg), off
        table[i]["numbps"], table[i]["numpasses"] = 1 + s, npass
        chunks += [cup, seg]
        off += len(cup) + len(seg)
    data = np.concatenate(chunks)
    cs = G.codestream_write(cp, table, data)
    cp2, blocks = G.codestream_parse(cs)
    for f in ("numbps", "numpasses", "length", "length2"):
        assert np.array_equal(blocks[f], table[f]), f
    assert (blocks["numpasses"] == 3).any() and (blocks["numpasses"] == 2).any()
    ref = P.inverse(cp, [want])[0]
    got = openjpeg_cv2(cs).astype(np.int64)
    assert np.array_equal(got, ref)


def test_parser_declines_or_rejects_what_it_does_not_cover():
    cp = G.make_coding(64, 64, 1, 8, numres=3)
    planes = P.synthetic_image(64, 64, 1, 8, seed=24)
    table, data, _ = oracle_encode(cp, planes)
    cs = G.codestream_write(cp, table, data)
    bad = cs.copy()
    bad[0] = 0                                   # no SOC
    with pytest.raises(G.EngineError):
        G.codestream_parse(bad)
    i = bytes(cs).find(b"\xff\x52")              # COD: two layers
    two = cs.copy()
    two[i + 6] = 2
    with pytest.raises(G.NotHandled):
        G.codestream_parse(two)
    sty = cs.copy()
    sty[i + 12] = 0                              # Part-1 block coder
    with pytest.raises(G.NotHandled):
        G.codestream_parse(sty)
    with pytest.raises(G.EngineError):
        G.codestream_parse(cs[:len(cs) // 2])    # truncated: packets run past the end
    rng = np.random.default_rng(5)
    for _ in range(200):                         # random damage never crashes the parser
        d = cs.copy()
        k = rng.integers(0, len(d), 4)
        d[k] ^= rng.integers(1, 256, 4).astype(np.uint8)
        try:
            G.codestream_parse(d)
        except G.EngineError:
            pass


def test_jph_container_round_trip_and_openjpeg():
    """.jph (JP2 boxes, brand 'jph '): OpenJPEG opens the wrapped file, and the codestream comes back out of it."""
    cp = G.make_coding(160, 96, 3, 8, numres=4, tile=(64, 64))
    planes = P.synthetic_image(160, 96, 3, 8, seed=41)
    table, data, _ = oracle_encode(cp, planes)
    cs = G.codestream_write(cp, table, data)
    jph = G.jph_wrap(cp, cs)
    assert bytes(jph[4:8]) == b"jP  " and bytes(jph[16:24]) == b"ftypjph "
    assert np.array_equal(G.jph_codestream(jph), cs)
    assert np.array_equal(G.jph_codestream(cs), cs)          # a raw codestream is accepted as it is
    got = openjpeg_pillow(jph)
    assert np.array_equal(got.astype(np.int64), np.stack(planes, axis=-1))
    with pytest.raises(G.EngineError):
        G.jph_codestream(jph[:40])


@pytest.mark.parametrize("prog", [G.LRCP, G.RLCP, G.RPCL, G.PCRL, G.CPRL])
@pytest.mark.parametrize("tparts", [0, G.CS_TPARTS_R])
def test_progression_orders_and_tile_parts(prog, tparts):
    """The five progression orders (position-driven ones with precincts of different sizes per resolution, a tile
    grid that does not start at the image origin, ragged tiles) and a tile part per resolution: OpenJPEG decodes
    every variant exactly, and the parser reads every variant back to the same block table and bytes."""
    args = dict(width=290, height=203, numcomps=3, prec=8, numres=4, tile=(128, 96), origin=(5, 3), tile_origin=(2, 1),
                precincts=[(16, 16), (32, 16), (32, 64), (64, 64)], cblk=(16, 16))
    cp = G.make_coding(**args)
    planes = P.synthetic_image(args["width"], args["height"], 3, 8, seed=51, origin=args["origin"])
    table, data, _ = oracle_encode(cp, planes)
    flags = G.CS_TLM | G.CS_PLT | tparts | G.CS_PROG(prog)
    cs = G.codestream_write(cp, table, data, flags)
    i = bytes(cs).find(b"\xff\x52")
    assert cs[i + 5] == prog
    nparts = bytes(cs).count(b"\xff\x90\x00\x0a")
    ntiles = len(P.tile_rects(cp))
    assert nparts == (ntiles * cp.numres if (tparts and prog <= G.RPCL) else ntiles)
    got = openjpeg_pillow(cs)
    assert np.array_equal(got.astype(np.int64), np.stack(planes, axis=-1))
    cp2, blocks = G.codestream_parse(cs)
    for f in ("numbps", "numpasses", "length"):
        assert np.array_equal(blocks[f], table[f]), f
    for k in range(0, len(table), 5):
        o, n = int(blocks[k]["offset"]), int(blocks[k]["length"])
        assert np.array_equal(cs[o:o + n], data[int(table[k]["offset"]):int(table[k]["offset"]) + n])


@pytest.mark.parametrize("irreversible", [False, True])
def test_explicit_qcd_exponents(irreversible):
    """b2k_coding.qcd_explicit: band exponents / mantissas other than the HT quantiser's (what a foreign encoder
    signals).  The product's geometry follows them (Kmax, step sizes), the writer puts them into QCD, OpenJPEG
    dequantises with them, and the parser hands them back."""
    w, h = 192, 160
    cp = G.make_coding(w, h, 3, 8, numres=4, irreversible=irreversible)
    base_e, base_m = P.quant_tables(cp)
    cp.qcd_explicit = 1
    for i in range(len(base_e)):
        cp.qcd_expn[i] = int(base_e[i]) + (1 if not irreversible else -1 + (i % 2))   # more head room / other step sizes
        cp.qcd_mant[i] = (int(base_m[i]) + 37 * i) % 2048 if irreversible else 0
    for gb in G.enumerate_blocks(cp):
        kmax, step_enc, _ = P.band_params(cp, int(gb["resno"]), int(gb["orient"]))
        assert gb["kmax"] == kmax and gb["stepsize"] == np.float32(step_enc)
    planes = P.synthetic_image(w, h, 3, 8, seed=61)
    table, data, _ = oracle_encode(cp, planes)
    cs = G.codestream_write(cp, table, data)
    cp2, blocks = G.codestream_parse(cs)
    n = len(base_e)
    assert cp2.qcd_explicit == 1 and list(cp2.qcd_expn)[:n] == list(cp.qcd_expn)[:n]
    if irreversible:
        assert list(cp2.qcd_mant)[:n] == list(cp.qcd_mant)[:n]
    got = openjpeg_cv2(cs).astype(np.int64)
    ours = np.stack(oracle_decode(cp2, blocks, cs), axis=-1).astype(np.int64)
    src = np.stack(planes, axis=-1).astype(np.int64)
    if irreversible:
        assert np.abs(got - ours).max() <= 1 and np.abs(ours - src).max() <= 6
    else:
        assert np.array_equal(got, src) and np.array_equal(ours, src)
    # the default tables come back as "not explicit"
    cp3 = G.make_coding(w, h, 3, 8, numres=4, irreversible=irreversible)
    t3, d3, _ = oracle_encode(cp3, planes)
    assert G.codestream_parse(G.codestream_write(cp3, t3, d3))[0].qcd_explicit == 0


@pytest.mark.parametrize("flags", [G.CS_SOP, G.CS_EPH, G.CS_SOP | G.CS_EPH | G.CS_PLT | G.CS_TPARTS_R])
def test_sop_and_eph_markers(flags):
    cp = G.make_coding(150, 100, 3, 8, numres=4, tile=(64, 64), precincts=[(32, 32)])
    planes = P.synthetic_image(150, 100, 3, 8, seed=52)
    table, data, _ = oracle_encode(cp, planes)
    cs = G.codestream_write(cp, table, data, flags)
    assert (bytes(cs).count(b"\xff\x91\x00\x04") > 0) == bool(flags & G.CS_SOP)
    got = openjpeg_pillow(cs)
    assert np.array_equal(got.astype(np.int64), np.stack(planes, axis=-1))
    cp2, blocks = G.codestream_parse(cs)
    assert np.array_equal(blocks["length"], table["length"]) and np.array_equal(blocks["numbps"], table["numbps"])
    for k in range(0, len(table), 3):
        o, n = int(blocks[k]["offset"]), int(blocks[k]["length"])
        assert np.array_equal(cs[o:o + n], data[int(table[k]["offset"]):int(table[k]["offset"]) + n])


def test_parser_tolerates_what_the_standard_allows():
    """Things other writers do: a COM segment in the main and in a tile-part header, Psot = 0 on the last tile part
    ("until EOC"), a missing EOC, no TLM / PLT, unknown informative marker segments."""
    cp = G.make_coding(130, 90, 1, 8, numres=3, tile=(64, 64))
    planes = P.synthetic_image(130, 90, 1, 8, seed=53)
    table, data, _ = oracle_encode(cp, planes)
    cs = bytes(G.codestream_write(cp, table, data, 0))

    def check(buf):
        buf = np.frombuffer(buf, np.uint8)
        cp2, blocks = G.codestream_parse(buf)
        assert np.array_equal(blocks["length"], table["length"])
        rec = oracle_decode(cp2, blocks, buf)
        assert np.array_equal(rec[0], planes[0])

    check(cs)
    com = b"\xff\x64" + (2 + 2 + 5).to_bytes(2, "big") + b"\x00\x01hello"
    i = cs.find(b"\xff\x52")                                    # before COD
    check(cs[:i] + com + cs[i:])
    # COM inside the first tile-part header: Psot grows by its length
    j = cs.find(b"\xff\x90")
    psot = int.from_bytes(cs[j + 6:j + 10], "big")
    tp = cs[j:j + 12][:6] + (psot + len(com)).to_bytes(4, "big") + cs[j + 10:j + 12]
    check(cs[:j] + tp + com + cs[j + 12:])
    # unknown informative segment (0xFF30-0xFF3F have no parameters and are not used here; take 0xFF70 with a length)
    unk = b"\xff\x70" + (2 + 3).to_bytes(2, "big") + b"abc"
    check(cs[:i] + unk + cs[i:])
    # last tile part with Psot = 0, with and without EOC
    k = cs.rfind(b"\xff\x90\x00\x0a")
    last0 = cs[:k + 6] + b"\x00\x00\x00\x00" + cs[k + 10:]
    check(last0)
    check(last0[:-2])
    check(cs[:-2])                                              # EOC missing, Psot intact


@pytest.mark.parametrize("args", CASES[:5] + [dict(width=290, height=203, numcomps=3, prec=8, numres=4, tile=(128, 96), origin=(5, 3),
                                                   tile_origin=(2, 1), precincts=[(16, 16), (32, 16), (32, 64), (64, 64)], cblk=(16, 16))])
def test_writer_matches_the_python_t2_oracle_byte_for_byte(args):
    """Two implementations of the T2 step, written independently (grok_b200/csrc/codestream.cpp in C++, tests/oracle_t2.py
    as a plain restatement of T2Compress.cpp / the marker writers): identical codestreams, byte for byte, with and
    without TLM + PLT -- and OpenJPEG decodes the oracle's stream as well."""
    import oracle_t2 as T2
    cp = G.make_coding(**args)
    planes = P.synthetic_image(args["width"], args["height"], args["numcomps"], args["prec"], seed=21,
                               origin=args.get("origin", (0, 0)))
    table, data, _ = oracle_encode(cp, planes)
    for tlm_plt in (False, True):
        want = T2.write_codestream(cp, table, data, tlm=tlm_plt, plt=tlm_plt)
        got = G.codestream_write(cp, table, data, (G.CS_TLM | G.CS_PLT) if tlm_plt else 0)
        assert len(want) == len(got) and np.array_equal(want, got)
    if args["prec"] == 8 or args["numcomps"] == 1:
        dec = openjpeg_pillow(want)
        src = planes[0] if len(planes) == 1 else np.stack(planes, axis=-1)
        assert np.array_equal(dec.astype(np.int64), src)


def test_parser_survives_mutated_streams():
    """Untrusted input: mutated / truncated code streams and JPH files (tests/fuzz_parser_driver.py, in a subprocess) always
    come back as an error, "not handled" or a block table inside the buffer.  (The same driver runs against an
    AddressSanitizer + UBSan build of codestream.cpp / geometry.cpp when one is passed to it; that is how the endless QCD loop
    pinned below was found.)"""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    p = subprocess.run([sys.executable, os.path.join(here, "fuzz_parser_driver.py"), "11", "300"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0 and "FUZZ ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_qcd_with_a_dangling_byte_is_an_error_not_a_loop():
    """A 16-bit-per-band QCD (style 1 / 2) whose payload has an odd byte left over used to spin forever, allocating."""
    args = dict(width=64, height=48, numcomps=1, prec=8, numres=3, irreversible=True)
    cp = G.make_coding(**args)
    planes = P.synthetic_image(64, 48, 1, 8, seed=3)
    table, data, _ = oracle_encode(cp, planes)
    cs = bytearray(G.codestream_write(cp, table, data, 0).tobytes())
    i = cs.index(b"\xff\x5c")                       # QCD
    lqcd = (cs[i + 2] << 8) | cs[i + 3]
    assert (cs[i + 4] & 0x1F) == 2                    # expounded: 16-bit entries
    cs[i + 2:i + 4] = bytes([(lqcd + 1) >> 8, (lqcd + 1) & 0xFF])
    cs.insert(i + 2 + lqcd, 0)                        # one dangling byte inside the segment
    with pytest.raises((G.EngineError, G.NotHandled)):
        G.codestream_parse(np.frombuffer(bytes(cs), np.uint8))
