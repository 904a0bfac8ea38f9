"""Minimal independent FITS writer/reader for the tests (numpy only; test infrastructure)."""
import numpy as np


def _card(key, value, comment=""):
    return ("%-8s= %20s / %-47s" % (key, value, comment))[:80].ljust(80)


def write_fits(path, data, bitpix=-32, bzero=None, bscale=None, exposure=None, extra=()):
    """data: 2-D array already in the on-disk type for integer BITPIX, float for -32/-64"""
    h, w = data.shape
    cards = [_card("SIMPLE", "T"), _card("BITPIX", bitpix), _card("NAXIS", 2), _card("NAXIS1", w), _card("NAXIS2", h)]
    if bzero is not None:
        cards.append(_card("BZERO", bzero))
    if bscale is not None:
        cards.append(_card("BSCALE", bscale))
    if exposure is not None:
        cards.append(_card("EXPTIME", exposure))
    cards += list(extra)
    cards.append("END".ljust(80))
    hdr = "".join(cards)
    hdr += " " * (-len(hdr) % 2880)
    dt = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}[bitpix]
    raw = np.ascontiguousarray(data, dtype=dt).tobytes()
    raw += b"\0" * (-len(raw) % 2880)
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii") + raw)


def read_fits(path):
    """-> (header dict of raw strings, float32 data as stored for BITPIX -32)"""
    with open(path, "rb") as f:
        blob = f.read()
    hdr, pos, end = {}, 0, False
    while not end:
        block = blob[pos:pos + 2880].decode("ascii")
        pos += 2880
        for i in range(36):
            line = block[80 * i:80 * i + 80]
            if line.startswith("END"):
                end = True
                break
            if "=" in line[:10]:
                hdr[line[:8].strip()] = line[10:].split("/")[0].strip()
    assert hdr["BITPIX"] == "-32"
    w, h = int(hdr["NAXIS1"]), int(hdr["NAXIS2"])
    data = np.frombuffer(blob[pos:pos + 4 * w * h], dtype=">f4").astype(np.float32).reshape(h, w)
    assert (len(blob) - pos) % 2880 == 0 and len(blob) - pos >= 4 * w * h
    return hdr, data
