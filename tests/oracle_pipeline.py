"""Whole-tile CPU pipeline assembled from the oracle's stages (TEST INFRASTRUCTURE ONLY).

Mirrors the order of the reference's per-tile DAG (TileProcessorCompress.cpp L347-531):
DC shift + MCT -> forward DWT per component -> HT cleanup coding of every code block, and the
inverse order for decode.  Used by the parity tests, smoke() and bench.py's cpu_baseline leg."""
import ctypes as C

import numpy as np

import oracle_lib as O


def tile_rects(cp):
    if cp.tw == 0:
        return [(cp.x0, cp.y0, cp.x1, cp.y1)]
    nx = -(-(cp.x1 - cp.tx0) // cp.tw)
    ny = -(-(cp.y1 - cp.ty0) // cp.th)
    out = []
    for q in range(ny):
        for p in range(nx):
            x0 = max(cp.tx0 + p * cp.tw, cp.x0)
            y0 = max(cp.ty0 + q * cp.th, cp.y0)
            x1 = min(cp.tx0 + (p + 1) * cp.tw, cp.x1)
            y1 = min(cp.ty0 + (q + 1) * cp.th, cp.y1)
            out.append((x0, y0, x1, y1))
    return out


def dc_shift(cp):
    return 0 if cp.sgnd else -(1 << (cp.prec - 1))


def forward(cp, planes, tiles=None):
    """planes: list of (H, W) int32.  Returns coefficient planes (int32; float bits when
    irreversible), Mallat layout inside each tile's rectangle."""
    L = O.lib()
    H, W = planes[0].shape
    out = [np.zeros((H, W), np.int32) for _ in planes]
    shift = np.array([dc_shift(cp)] * 3, np.int32)
    rects = tile_rects(cp)
    for t, (x0, y0, x1, y1) in enumerate(rects):
        if tiles is not None and t not in tiles:
            continue
        w, h = x1 - x0, y1 - y0
        sl = (slice(y0 - cp.y0, y1 - cp.y0), slice(x0 - cp.x0, x1 - cp.x0))
        comps = [np.ascontiguousarray(p[sl]).astype(np.int32) for p in planes]
        if cp.irreversible:
            fl = [None] * len(comps)
            if cp.mct:
                y, u, v = (np.zeros(w * h, np.float32) for _ in range(3))
                L.orc_ict_fwd(comps[0].ravel(), comps[1].ravel(), comps[2].ravel(), y, u, v, w * h, shift)
                fl[0], fl[1], fl[2] = y.reshape(h, w), u.reshape(h, w), v.reshape(h, w)
            for c in range(len(comps)):
                if fl[c] is None:
                    fl[c] = (comps[c] + shift[0]).astype(np.float32)
                buf = np.ascontiguousarray(fl[c])
                L.orc_dwt97_fwd_2d(buf, w, x0, y0, x1, y1, cp.numres)
                out[c][sl] = buf.view(np.int32)
        else:
            if cp.mct:
                a, b, c2 = (np.ascontiguousarray(k).ravel() for k in comps[:3])
                L.orc_rct_fwd(a, b, c2, w * h, shift)
                comps[0], comps[1], comps[2] = a.reshape(h, w), b.reshape(h, w), c2.reshape(h, w)
            for c in range(len(comps)):
                buf = np.ascontiguousarray(comps[c] if (cp.mct and c < 3) else comps[c] + shift[0])
                L.orc_dwt53_fwd_2d(buf, w, x0, y0, x1, y1, cp.numres)
                out[c][sl] = buf
    return out


def inverse(cp, coefs, tiles=None):
    """Coefficient planes -> sample planes (clamped, DC shift restored)."""
    L = O.lib()
    H, W = coefs[0].shape
    out = [np.zeros((H, W), np.int32) for _ in coefs]
    sh = -dc_shift(cp)
    shift = np.array([sh] * 3, np.int32)
    lo = np.array([-(1 << (cp.prec - 1)) if cp.sgnd else 0] * 3, np.int32)
    hi = np.array([(1 << (cp.prec - 1)) - 1 if cp.sgnd else (1 << cp.prec) - 1] * 3, np.int32)
    for t, (x0, y0, x1, y1) in enumerate(tile_rects(cp)):
        if tiles is not None and t not in tiles:
            continue
        w, h = x1 - x0, y1 - y0
        sl = (slice(y0 - cp.y0, y1 - cp.y0), slice(x0 - cp.x0, x1 - cp.x0))
        if cp.irreversible:
            fl = []
            for c in range(len(coefs)):
                buf = np.ascontiguousarray(coefs[c][sl]).view(np.float32).copy()
                L.orc_dwt97_inv_2d(buf, w, x0, y0, x1, y1, cp.numres)
                fl.append(buf)
            if cp.mct:
                r, g, b = (np.zeros(w * h, np.int32) for _ in range(3))
                L.orc_ict_inv(fl[0].ravel(), fl[1].ravel(), fl[2].ravel(), r, g, b, w * h, shift, lo, hi)
                out[0][sl], out[1][sl], out[2][sl] = r.reshape(h, w), g.reshape(h, w), b.reshape(h, w)
            for c in range(3 if cp.mct else 0, len(coefs)):
                out[c][sl] = np.clip(np.rint(fl[c]).astype(np.int64) + sh, lo[0], hi[0]).astype(np.int32)
        else:
            comps = []
            for c in range(len(coefs)):
                buf = np.ascontiguousarray(coefs[c][sl]).copy()
                L.orc_dwt53_inv_2d(buf, w, x0, y0, x1, y1, cp.numres)
                comps.append(buf)
            if cp.mct:
                a, b, c2 = (np.ascontiguousarray(k).ravel() for k in comps[:3])
                L.orc_rct_inv(a, b, c2, w * h, shift, lo, hi)
                out[0][sl], out[1][sl], out[2][sl] = a.reshape(h, w), b.reshape(h, w), c2.reshape(h, w)
            for c in range(3 if cp.mct else 0, len(coefs)):
                out[c][sl] = np.clip(comps[c].astype(np.int64) + sh, lo[0], hi[0]).astype(np.int32)
    return out


def quant_tables(cp):
    n = 3 * (cp.numres - 1) + 1
    expn = np.zeros(n, np.uint8)
    mant = np.zeros(n, np.uint16)
    got = O.lib().orc_ht_stepsizes(cp.numres - 1, cp.prec, cp.mct, cp.sgnd, 0 if cp.irreversible else 1, expn, mant)
    assert got == n
    if getattr(cp, "qcd_explicit", 0):        # a foreign stream's QCD / the caller's own exponents (b2k_coding)
        expn = np.array(list(cp.qcd_expn)[:n], np.uint8)
        mant = np.array(list(cp.qcd_mant)[:n], np.uint16) if cp.irreversible else np.zeros(n, np.uint16)
    return expn, mant


def band_params(cp, resno, orient):
    expn, mant = quant_tables(cp)
    i = 0 if resno == 0 else 1 + 3 * (resno - 1) + (orient - 1)
    L = O.lib()
    kmax = L.orc_band_kmax(int(expn[i]), cp.numgbits, 0)
    step_enc = L.orc_band_stepsize(cp.prec, orient, int(expn[i]), int(mant[i]), 1, 0 if cp.irreversible else 1)
    step_dec = L.orc_band_stepsize(cp.prec, orient, int(expn[i]), int(mant[i]), 0, 0 if cp.irreversible else 1)
    return kmax, step_enc, step_dec


def enumerate_all(cp, tiles=None):
    """[(tile, comp, oracle Block)] in the reference's order tile->comp->res->band->prec->cblk."""
    out = []
    for t, (x0, y0, x1, y1) in enumerate(tile_rects(cp)):
        if tiles is not None and t not in tiles:
            continue
        blks = O.enumerate_blocks((x0, y0, x1, y1), cp.numres, cp.cblkw_exp, cp.cblkh_exp, list(cp.prcw_exp), list(cp.prch_exp))
        for c in range(cp.numcomps):
            for b in blks:
                out.append((t, c, b))
    return out


def encode_block(cp, coefs, rect, comp, b):
    """HT-encode one code block straight from the coefficient planes; returns bytes (uint8)."""
    L = O.lib()
    x0, y0 = rect[0] - cp.x0, rect[1] - cp.y0
    w, h = b.x1 - b.x0, b.y1 - b.y0
    kmax, step_enc, _ = band_params(cp, b.resno, b.orient)
    win = np.ascontiguousarray(coefs[comp][y0 + b.buf_y:y0 + b.buf_y + h, x0 + b.buf_x:x0 + b.buf_x + w])
    sm = np.zeros(w * h, np.uint32)
    if cp.irreversible:
        L.orc_ht_pre_irrev(win.view(np.float32), w, w, h, kmax, np.float32(1.0) / np.float32(step_enc), sm)
    else:
        L.orc_ht_pre_rev(win, w, w, h, kmax, sm)
    return O.ht_encode(sm.reshape(h, w), kmax)


def decode_block(cp, data, comp, b, numbps=1):
    """Bytes -> dequantised coefficient window (int32 / float bits), PostDecodeFiltersOJPH.h."""
    L = O.lib()
    w, h = b.x1 - b.x0, b.y1 - b.y0
    kmax, _, step_dec = band_params(cp, b.resno, b.orient)
    if len(data) == 0:
        return np.zeros((h, w), np.int32)
    rc, dec = O.ht_decode(data, kmax - numbps, w, h)
    assert rc == 0
    if cp.irreversible:
        out = np.zeros((h, w), np.float32)
        L.orc_ht_post_irrev(dec, w, w, h, kmax, step_dec, out, w)
        return out.view(np.int32)
    out = np.zeros((h, w), np.int32)
    L.orc_ht_post_rev(dec, w, w, h, kmax, out, w)
    return out


def synthetic_image(width, height, ncomp, prec, seed, origin=(0, 0)):
    """SURVEY.md section 8d generator: smooth gradient + sinusoid + 5 bits of noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[origin[1]:origin[1] + height, origin[0]:origin[0] + width].astype(np.int64)
    planes = []
    scale = max(1, (1 << prec) // 4096)
    for c in range(ncomp):
        base = (x * (3 + c) + y * (5 - c)) // 16 + (64 * np.sin((x + 2 * y) / (97.0 + 13 * c))).astype(np.int64)
        noise = rng.integers(0, 32, size=(height, width))
        planes.append((((base + noise) * scale) % (1 << prec)).astype(np.int32))
    return planes
