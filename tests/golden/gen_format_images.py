"""Test images for the host's image ingest: one small file per flavour of JPEG / BMP / GIF / TGA / PSD / 16-bit PNM
(PIL where it can write the flavour, by hand otherwise).  Called by make_golden_cli.py, which then records what the
reference's stb_image decodes from each of them.

    python tests/golden/gen_format_images.py tests/golden/images
"""
import struct
import sys
from pathlib import Path

import numpy as np
from PIL import Image

out = Path(sys.argv[1])
out.mkdir(exist_ok=True)
rng = np.random.RandomState(21)
w, h = 19, 13


def smooth(hh, ww, c):
    base = rng.randint(0, 256, size=(hh // 6 + 2, ww // 6 + 2, c)).astype(np.float64)
    up = np.kron(base, np.ones((6, 6, 1)))[:hh, :ww]
    return np.clip(up + rng.randint(-25, 26, size=(hh, ww, c)), 0, 255).astype(np.uint8)


rgb = smooth(h, w, 3)
rgba = np.dstack([rgb, rng.randint(0, 256, size=(h, w)).astype(np.uint8)])
rgba[::3, ::2, 3] = 0
grey = smooth(h, w, 1)[..., 0]

# ---- JPEG ----
big = smooth(67, 90, 3)
Image.fromarray(rgb, "RGB").save(out / "base_444.jpg", quality=90, subsampling=0)
Image.fromarray(big, "RGB").save(out / "base_422.jpg", quality=85, subsampling=1)
Image.fromarray(big, "RGB").save(out / "base_420.jpg", quality=75, subsampling=2)
Image.fromarray(big, "RGB").save(out / "base_420_restart.jpg", quality=75, subsampling=2, restart_marker_blocks=2)
Image.fromarray(big, "RGB").save(out / "prog_420.jpg", quality=80, subsampling=2, progressive=True)
Image.fromarray(rgb, "RGB").save(out / "prog_444.jpg", quality=95, subsampling=0, progressive=True)
Image.fromarray(grey, "L").save(out / "grey.jpg", quality=80)
Image.fromarray(grey, "L").save(out / "grey_prog.jpg", quality=60, progressive=True)
Image.fromarray(rgb, "RGB").save(out / "huff_opt.jpg", quality=50, optimize=True)
Image.fromarray(smooth(h, w, 4), "CMYK").save(out / "cmyk.jpg", quality=85)
Image.fromarray(rgb[:1, :1], "RGB").save(out / "one_pixel.jpg")

# ---- BMP ----
Image.fromarray(rgb, "RGB").save(out / "rgb24.bmp")
Image.fromarray(rgba, "RGBA").save(out / "rgba32.bmp")
pal = Image.fromarray(rgb, "RGB").quantize(colors=200)
pal.save(out / "pal8.bmp")
Image.fromarray(rgb, "RGB").quantize(colors=13).save(out / "pal4.bmp", bits=4)
Image.fromarray((grey > 128).astype(np.uint8) * 255, "L").convert("1").save(out / "mono1.bmp")


def bmp(path, width, height, bpp, rows, compress=0, masks=None, hsz=40, palette=None, top_down=False):
    """rows: list of bytes per row (unpadded), top row first."""
    pad = (-len(rows[0])) & 3
    data = b"".join(r + b"\0" * pad for r in (rows if top_down else rows[::-1]))
    ph = b""
    if hsz == 12:
        hdr = struct.pack("<IHHHH", 12, width, height, 1, bpp)
    else:
        hdr = struct.pack("<IiiHHIIiiII", hsz, width, -height if top_down else height, 1, bpp, compress, len(data), 2835, 2835,
                          0, 0)
        if hsz in (108, 124):
            m = masks or (0, 0, 0, 0)
            hdr += struct.pack("<IIII", *m) + b"\0" * (4 + 48) + (b"\0" * 16 if hsz == 124 else b"")
        elif hsz == 56:
            hdr += struct.pack("<IIII", *(masks or (0, 0, 0, 0)))
        elif compress == 3:
            ph = struct.pack("<III", *masks[:3])
    if palette is not None:
        ph += palette
    off = 14 + len(hdr) + len(ph)
    with open(path, "wb") as f:
        f.write(b"BM" + struct.pack("<IHHI", off + len(data), 0, 0, off) + hdr + ph + data)


px565 = ((rgb[..., 0].astype(np.uint16) >> 3) << 11) | ((rgb[..., 1].astype(np.uint16) >> 2) << 5) | (rgb[..., 2] >> 3)
px555 = ((rgb[..., 0].astype(np.uint16) >> 3) << 10) | ((rgb[..., 1].astype(np.uint16) >> 3) << 5) | (rgb[..., 2] >> 3)
bmp(out / "rgb565_bitfields.bmp", w, h, 16, [r.astype("<u2").tobytes() for r in px565], compress=3, masks=(0xF800, 0x07E0, 0x001F))
bmp(out / "rgb555.bmp", w, h, 16, [r.astype("<u2").tobytes() for r in px555])
bgra = rgba[..., [2, 1, 0, 3]]
bmp(out / "bgra32_v4_bitfields.bmp", w, h, 32, [r.tobytes() for r in bgra], compress=3, hsz=108,
    masks=(0x00FF0000, 0x0000FF00, 0x000000FF, 0xFF000000))
bmp(out / "bgra32_v5.bmp", w, h, 32, [r.tobytes() for r in bgra], hsz=124, masks=(0x00FF0000, 0x0000FF00, 0x000000FF, 0xFF000000))
zero_a = bgra.copy()
zero_a[..., 3] = 0
bmp(out / "bgrx32_alpha_all_zero.bmp", w, h, 32, [r.tobytes() for r in zero_a])
bmp(out / "rgb24_topdown.bmp", w, h, 24, [r[:, ::-1].tobytes() for r in rgb], top_down=True)
xrgb = ((rgba[..., 3].astype(np.uint32) >> 4) << 12) | ((rgb[..., 0].astype(np.uint32) >> 4) << 8) | ((rgb[..., 1].astype(np.uint32) >> 4) << 4) | (rgb[..., 2] >> 4)
bmp(out / "argb4444_v4.bmp", w, h, 16, [r.astype("<u2").tobytes() for r in xrgb], compress=3, hsz=108, masks=(0x0F00, 0x00F0, 0x000F, 0xF000))
# OS/2 header: stb sizes the palette as (offset - 14 - 24) / 3 = 252 entries of the 256 stored here (a stb quirk);
# indices stay below that, beyond it stb reads uninitialised memory
idx = rng.randint(0, 252, size=(h, w)).astype(np.uint8)
bmp(out / "pal8_os2.bmp", w, h, 8, [r.tobytes() for r in idx], hsz=12, palette=bytes(rng.randint(0, 256, size=256 * 3).tolist()))

# ---- GIF ----
Image.fromarray(rgb, "RGB").quantize(colors=40).save(out / "plain.gif")
Image.fromarray(big, "RGB").quantize(colors=256).save(out / "interlaced.gif", interlace=True)
gp = Image.fromarray(rgb, "RGB").quantize(colors=17)
gp.save(out / "transparent.gif", transparency=3)
gp.save(out / "background.gif", background=5)


def gif_subimage(path):
    """A frame smaller than the screen, local colour table, background index 2: exercises the undrawn-pixel fill."""
    sw, sh, fw, fh = 12, 9, 5, 4
    gct = bytes(rng.randint(0, 256, size=4 * 3).tolist())
    lct = bytes(rng.randint(0, 256, size=4 * 3).tolist())
    pix = rng.randint(0, 4, size=fw * fh).tolist()
    # uncompressed-style LZW: code size 2 -> clear = 4, eoi = 5; emit a clear before every 2 pixels to keep 3-bit codes
    codes = []
    for i, p in enumerate(pix):
        if i % 2 == 0:
            codes.append(4)
        codes.append(p)
    codes.append(5)
    bits, nb, data = 0, 0, bytearray()
    for c in codes:
        bits |= c << nb
        nb += 3
        while nb >= 8:
            data.append(bits & 255)
            bits >>= 8
            nb -= 8
    if nb:
        data.append(bits & 255)
    with open(path, "wb") as f:
        f.write(b"GIF89a" + struct.pack("<HHBBB", sw, sh, 0x80 | 1, 2, 0) + gct)
        f.write(b"\x21\xF9\x04" + struct.pack("<BHB", 1, 0, 1) + b"\0")          # GCE: transparent index 1
        f.write(b"\x2C" + struct.pack("<HHHHB", 3, 2, fw, fh, 0x80 | 1) + lct)
        f.write(bytes([2, len(data)]) + bytes(data) + b"\0\x3B")


gif_subimage(out / "subimage_local_table.gif")

# ---- TGA ----
Image.fromarray(rgb, "RGB").save(out / "rgb24_rle.tga", compression="tga_rle")
Image.fromarray(rgba, "RGBA").save(out / "rgba32_rle.tga", compression="tga_rle")
Image.fromarray(grey, "L").save(out / "grey8.tga")
Image.fromarray(rgb, "RGB").quantize(colors=50).save(out / "pal8.tga")
Image.fromarray(np.dstack([grey, rgba[..., 3]]), "LA").save(out / "grey_alpha16.tga")


def tga(path, image_type, bpp, data, width=w, height=h, desc=0, cmap=None, cmap_bits=0, id_field=b""):
    hdr = struct.pack("<BBBHHBHHHHBB", len(id_field), 1 if cmap is not None else 0, image_type, 0,
                      (len(cmap) * 8 // cmap_bits) if cmap is not None else 0, cmap_bits, 0, 0, width, height, bpp, desc)
    with open(path, "wb") as f:
        f.write(hdr + id_field + (cmap or b"") + data)


tga(out / "rgb16_topdown_noext", 2, 16, px555.astype("<u2").tobytes(), desc=0x20, id_field=b"hello")   # sniffed, no .tga suffix
tga(out / "rgb15.tga", 2, 15, px555[::-1].astype("<u2").tobytes())
pal16 = rng.randint(0, 32768, size=7).astype("<u2")
tga(out / "pal16_idx8.tga", 1, 8, rng.randint(0, 7, size=w * h).astype(np.uint8).tobytes(), cmap=pal16.tobytes(), cmap_bits=16)
pal32 = rng.randint(0, 256, size=300 * 4).astype(np.uint8)
tga(out / "pal32_idx16.tga", 1, 16, rng.randint(0, 300, size=w * h).astype("<u2").tobytes(), cmap=pal32.tobytes(), cmap_bits=32)

# ---- PSD ----


def psd(path, channels, depth, planes, rle=False):
    hh, ww = planes[0].shape
    with open(path, "wb") as f:
        f.write(b"8BPS" + struct.pack(">H6xHIIHH", 1, channels, hh, ww, depth, 3))
        f.write(struct.pack(">I", 0) + struct.pack(">I", 4) + b"abcd" + struct.pack(">I", 0))
        f.write(struct.pack(">H", 1 if rle else 0))
        if not rle:
            for p in planes:
                f.write(p.astype(">u2").tobytes() if depth == 16 else p.astype(np.uint8).tobytes())
            return
        rows = []
        for p in planes:
            for r in p:
                enc, i = bytearray(), 0
                r = r.tolist()
                while i < len(r):
                    j = i
                    while j + 1 < len(r) and r[j + 1] == r[i] and j - i < 126:
                        j += 1
                    if j > i:
                        enc += bytes([257 - (j - i + 1), r[i]])
                        i = j + 1
                    else:
                        k = i
                        while k + 1 < len(r) and r[k + 1] != r[k] and k - i < 126:
                            k += 1
                        enc += bytes([k - i]) + bytes(r[i:k + 1])
                        i = k + 1
                rows.append(bytes(enc))
        f.write(b"".join(struct.pack(">H", len(r)) for r in rows) + b"".join(rows))


flat = (rgb // 32) * 32
psd(out / "rgb8_raw.psd", 3, 8, [rgb[..., c] for c in range(3)])
psd(out / "rgba8_rle.psd", 4, 8, [flat[..., 0], flat[..., 1], flat[..., 2], (rgba[..., 3] // 64) * 64 + 63], rle=True)
psd(out / "rgba16_raw.psd", 4, 16, [(rgba[..., c].astype(np.uint16) * 257) for c in range(4)])
psd(out / "grey_as_rgb_1ch.psd", 1, 8, [grey])

# ---- PNM 16 bit ----
v16 = rng.randint(0, 65536, size=(h, w, 3)).astype(">u2")
(out / "rgb16.ppm").write_bytes(b"P6\n%d %d\n65535\n" % (w, h) + v16.tobytes())
# ---- Radiance HDR (RGBE) ----
rng2 = np.random.RandomState(77)


def rgbe_pixels(hh, ww):
    """Random RGBE quadruples: exponents around 128 (values around 0.1 .. 4), some zero exponents, some saturating."""
    px = rng2.randint(0, 256, size=(hh, ww, 4)).astype(np.uint8)
    px[..., 3] = rng2.randint(120, 132, size=(hh, ww))
    px[::4, ::3, 3] = 0              # exponent 0: black whatever the mantissas say
    px[1::5, 1::4, 3] = 140          # far above 1.0: clamps to 255
    px[..., 0] |= 0x80               # a normalised pixel has one mantissa >= 128: never mistaken for an RLE marker
    return px


def hdr_rle_rows(px):
    """The 'new' run-length scanlines: 2 2 hi lo, then each of the four components run-length coded."""
    data = b""
    for row in px:
        data += bytes([2, 2, row.shape[0] >> 8, row.shape[0] & 255])
        for k in range(4):
            comp, i = row[:, k].tolist(), 0
            while i < len(comp):
                j = i
                while j + 1 < len(comp) and comp[j + 1] == comp[i] and j - i < 126:
                    j += 1
                if j - i >= 2:
                    data += bytes([128 + (j - i + 1), comp[i]])
                    i = j + 1
                else:
                    k2 = i
                    while k2 + 1 < len(comp) and k2 - i < 127 and not (k2 + 2 < len(comp) and comp[k2 + 1] == comp[k2 + 2]):
                        k2 += 1
                    data += bytes([k2 - i + 1]) + bytes(comp[i:k2 + 1])
                    i = k2 + 1
    return data


def hdr(path, px, body, magic=b"#?RADIANCE", extra=b"# made by hand\nEXPOSURE=1.0\n"):
    hh, ww = px.shape[:2]
    path.write_bytes(magic + b"\n" + extra + b"FORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (hh, ww) + body)


e = rgbe_pixels(h, w)
e[:, 5:12, :3] = e[:, 5:6, :3]       # runs
hdr(out / "rle.hdr", e, hdr_rle_rows(e))
hdr(out / "rgbe_magic_rle.hdr", e[:6], hdr_rle_rows(e[:6]), magic=b"#?RGBE", extra=b"")
hdr(out / "flat_wide.hdr", e, e.tobytes())                       # width >= 8 but stored flat
small = rgbe_pixels(9, 5)
hdr(out / "flat_narrow.hdr", small, small.tobytes())             # width < 8: always flat
# a run-length first scanline followed by flat data: the reference restarts at pixel 0 with those four bytes
mixed = hdr_rle_rows(e[:1]) + np.ascontiguousarray(e[::-1]).tobytes()     # (so the image is e upside down)
hdr(out / "rle_then_flat.hdr", e, mixed)

# ---- Softimage PIC ----


def pic(path, ww, hh, packets, rows):
    """packets: [(type, channel mask)], rows: per scanline the bytes of each packet in turn."""
    head = b"\x53\x80\xF6\x34" + struct.pack(">f", 3.71) + b"hand made".ljust(80, b"\0") + b"PICT"
    head += struct.pack(">HHfHH", ww, hh, 1.0, 3, 0)
    for i, (t, c) in enumerate(packets):
        head += bytes([1 if i + 1 < len(packets) else 0, 8, t, c])
    path.write_bytes(head + b"".join(rows))


def pic_mixed(vals):
    """Mixed run-length coding of a list of per-pixel byte tuples."""
    data, i = b"", 0
    while i < len(vals):
        j = i
        while j + 1 < len(vals) and vals[j + 1] == vals[i]:
            j += 1
        n = j - i + 1
        if n >= 130:
            data += bytes([128]) + struct.pack(">H", n) + bytes(vals[i])
            i = j + 1
        elif n >= 2:
            data += bytes([n + 127]) + bytes(vals[i])
            i = j + 1
        else:
            k2 = i
            while k2 + 1 < len(vals) and k2 - i < 127 and not (k2 + 2 < len(vals) and vals[k2 + 1] == vals[k2 + 2]):
                k2 += 1
            data += bytes([k2 - i]) + b"".join(bytes(v) for v in vals[i:k2 + 1])
            i = k2 + 1
    return data


pic(out / "rgb_raw.pic", w, h, [(0, 0xE0)], [r.tobytes() for r in rgb])
pic(out / "rgb_plus_alpha_raw.pic", w, h, [(0, 0xE0), (0, 0x10)], [rgba[y, :, :3].tobytes() + rgba[y, :, 3].tobytes() for y in range(h)])
wide = np.repeat(flat[:, :, :], 16, axis=1)[:, :300]             # long runs: the 16-bit repeat count
wide[:, 40:45] = rng2.randint(0, 256, size=(h, 5, 3))
pic(out / "rgb_mixed_rle.pic", 300, h, [(2, 0xE0)], [pic_mixed([tuple(p) for p in r.tolist()]) for r in wide])
row_pure = [b"".join(bytes([4]) + bytes(r[x0].tolist()) for x0 in range(0, 20, 4)) for r in rgb]      # last run clipped at the width
pic(out / "rgb_pure_rle.pic", w, h, [(1, 0xE0)], row_pure)
pic(out / "red_blue_only_mixed.pic", w, h, [(2, 0xA0)],
    [pic_mixed([(p[0], p[2]) for p in r.tolist()]) for r in flat])
print(len(list(out.iterdir())), "images")
