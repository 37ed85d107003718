"""Drop-in for the reference's lib/dsm_util.py (:38-159) WITHOUT GDAL: a small GeoTIFF writer/reader.

Same function names, arguments and return values:
  parse_proj_str(proj_str) -> (zone_number, hemisphere)
  read_dsm_tif(file) -> (float32 image with nodata -> NaN, meta_dict with the keys of :76-99)
  write_dsm_tif(image, out_file, geo, utm_zone, nodata_val=None)
  get_driver(file) -> an object with the two driver facts the reference uses (kept for API parity)

Files written: baseline little-endian TIFF, one float32 band, uncompressed strips, GeoTIFF keys for
EPSG:326xx/327xx (WGS 84 / UTM zone NN N|S), ModelPixelScale + ModelTiepoint equivalent to the reference's
geotransform (ul_e, res, 0, ul_n, 0, -res), RasterPixelIsArea (the reference's AREA_OR_POINT=Area) and the
GDAL_NODATA tag, so GDAL/QGIS read them like the reference's outputs.  The reader takes single-band float32 classic
TIFFs of either byte order, organised in strips or tiles, uncompressed or with the compressions GDAL users commonly
choose (Deflate, LZW, PackBits) and predictors 1, 2 and 3 -- so per-view DSMs written by the reference itself (GDAL
defaults or creation options) can be fed to the fusion.  BigTIFF and multi-band files are rejected.
"""
import os
import struct
import zlib

import numpy as np

# TIFF tags
_T_WIDTH, _T_LENGTH, _T_BITS, _T_COMPRESSION, _T_PHOTOMETRIC = 256, 257, 258, 259, 262
_T_STRIP_OFFSETS, _T_SPP, _T_ROWS_PER_STRIP, _T_STRIP_BYTES, _T_PLANAR, _T_SAMPLE_FORMAT = 273, 277, 278, 279, 284, 339
_T_PIXEL_SCALE, _T_TIEPOINT, _T_TRANSFORM, _T_GEOKEYS, _T_GEO_DOUBLES, _T_GEO_ASCII = 33550, 33922, 34264, 34735, 34736, 34737
_T_PREDICTOR, _T_TILE_WIDTH, _T_TILE_LENGTH, _T_TILE_OFFSETS, _T_TILE_BYTES = 317, 322, 323, 324, 325
_T_GDAL_METADATA, _T_GDAL_NODATA = 42112, 42113
_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8}
_TYPE_FMT = {1: 'B', 2: 'c', 3: 'H', 4: 'I', 6: 'b', 8: 'h', 9: 'i', 11: 'f', 12: 'd', 16: 'Q'}


def parse_proj_str(proj_str):
    idx1 = proj_str.find('UTM zone')
    idx2 = proj_str.find('",')
    sub_str = proj_str[idx1:idx2]
    hemisphere = sub_str[-1]
    zone_number = int(sub_str[-3:-1])
    return zone_number, hemisphere


def _utm_wkt(zone_number, hemisphere):
    """WKT in the shape GDAL reports for these files; parse_proj_str() works on it."""
    epsg = (32600 if hemisphere == 'N' else 32700) + int(zone_number)
    return ('PROJCS["WGS 84 / UTM zone {z}{h}",GEOGCS["WGS 84",DATUM["WGS_1984",SPHEROID["WGS 84",6378137,'
            '298.257223563,AUTHORITY["EPSG","7030"]],AUTHORITY["EPSG","6326"]],PRIMEM["Greenwich",0],'
            'UNIT["degree",0.0174532925199433],AUTHORITY["EPSG","4326"]],PROJECTION["Transverse_Mercator"],'
            'PARAMETER["latitude_of_origin",0],PARAMETER["central_meridian",{cm}],PARAMETER["scale_factor",0.9996],'
            'PARAMETER["false_easting",500000],PARAMETER["false_northing",{fn}],UNIT["metre",1],'
            'AXIS["Easting",EAST],AXIS["Northing",NORTH],AUTHORITY["EPSG","{e}"]]').format(
                z=int(zone_number), h=hemisphere, cm=int(zone_number) * 6 - 183,
                fn=0 if hemisphere == 'N' else 10000000, e=epsg)


class _Driver:
    ShortName = 'GTiff'

    def GetMetadataItem(self, key):
        return {'DCAP_RASTER': 'YES', 'DMD_EXTENSIONS': 'tif tiff'}.get(key)


def get_driver(file):
    f_ext = os.path.splitext(file)[1]
    return _Driver() if f_ext in ('.tif', '.tiff') else None


# out_file: .tif file to write
# geo: (ul_e, ul_n, e_resolution, n_resolution)
# utm_zone: (zone number, N or S)
def write_dsm_tif(image, out_file, geo, utm_zone, nodata_val=None):
    assert (len(image.shape) == 2)      # image should only be 2D
    ul_e, ul_n, e_resolution, n_resolution = geo
    zone_number, hemisphere = utm_zone

    # replace nan with no_data (:128-133)
    if nodata_val is not None:
        image = image.copy()    # avoid modify source data
        image[np.isnan(image)] = nodata_val
    else:
        nodata_val = np.nan
    data = np.ascontiguousarray(image.astype(np.float32)).astype('<f4', copy=False)
    height, width = data.shape
    if get_driver(out_file) is None:
        raise ValueError('no raster driver for {}'.format(out_file))

    epsg = (32600 if hemisphere == 'N' else 32700) + int(zone_number)
    citation = 'WGS84 / UTM zone {}{}|'.format(zone_number, hemisphere).encode('ascii')     # :153 SetProjCS name
    geokeys = [1, 1, 0, 4,
               1024, 0, 1, 1,            # GTModelTypeGeoKey = ModelTypeProjected
               1025, 0, 1, 1,            # GTRasterTypeGeoKey = RasterPixelIsArea (AREA_OR_POINT=Area, :157)
               1026, _T_GEO_ASCII, len(citation), 0,   # GTCitationGeoKey
               3072, 0, 1, epsg]         # ProjectedCSTypeGeoKey
    nodata_str = (repr(float(nodata_val)) if not float(nodata_val).is_integer() else str(int(nodata_val)))
    nodata_ascii = (('nan' if np.isnan(nodata_val) else nodata_str) + '\0').encode('ascii')
    gdal_meta = ('<GDALMetadata>\n  <Item name="AREA_OR_POINT">Area</Item>\n</GDALMetadata>\n\0').encode('ascii')

    rows_per_strip = max(1, min(height, (1 << 20) // max(width * 4, 1)))
    n_strips = (height + rows_per_strip - 1) // rows_per_strip
    strip_bytes = [min(rows_per_strip, height - i * rows_per_strip) * width * 4 for i in range(n_strips)]

    entries = []        # (tag, type, count, payload bytes)

    def add(tag, typ, values):
        if typ == 2:
            payload = values
            count = len(values)
        else:
            payload = struct.pack('<{}{}'.format(len(values), _TYPE_FMT[typ]), *values)
            count = len(values)
        entries.append((tag, typ, count, payload))

    add(_T_WIDTH, 4, [width])
    add(_T_LENGTH, 4, [height])
    add(_T_BITS, 3, [32])
    add(_T_COMPRESSION, 3, [1])
    add(_T_PHOTOMETRIC, 3, [1])
    add(_T_STRIP_OFFSETS, 4, [0] * n_strips)            # patched below
    add(_T_SPP, 3, [1])
    add(_T_ROWS_PER_STRIP, 4, [rows_per_strip])
    add(_T_STRIP_BYTES, 4, strip_bytes)
    add(_T_PLANAR, 3, [1])
    add(_T_SAMPLE_FORMAT, 3, [3])
    add(_T_PIXEL_SCALE, 12, [float(e_resolution), float(n_resolution), 0.0])       # geotransform[1], -geotransform[5]
    add(_T_TIEPOINT, 12, [0.0, 0.0, 0.0, float(ul_e), float(ul_n), 0.0])           # geotransform[0], [3]
    add(_T_GEOKEYS, 3, geokeys)
    add(_T_GEO_ASCII, 2, citation + b'\0')
    add(_T_GDAL_METADATA, 2, gdal_meta)
    add(_T_GDAL_NODATA, 2, nodata_ascii)
    entries.sort(key=lambda e: e[0])

    ifd_offset = 8
    ifd_size = 2 + 12 * len(entries) + 4
    extra_offset = ifd_offset + ifd_size
    extra = b''
    placed = []
    for tag, typ, count, payload in entries:
        if len(payload) <= 4:
            placed.append((tag, typ, count, payload.ljust(4, b'\0'), None))
        else:
            if len(extra) % 2:
                extra += b'\0'
            placed.append((tag, typ, count, None, extra_offset + len(extra)))
            extra += payload
    data_offset = extra_offset + len(extra)
    data_offset += (-data_offset) % 16
    offsets = [data_offset + sum(strip_bytes[:i]) for i in range(n_strips)]
    off_payload = struct.pack('<{}I'.format(n_strips), *offsets)

    with open(out_file, 'wb') as fp:
        fp.write(b'II' + struct.pack('<HI', 42, ifd_offset))
        fp.write(struct.pack('<H', len(placed)))
        extra_patch = bytearray(extra)
        for tag, typ, count, inline, off in placed:
            if tag == _T_STRIP_OFFSETS:
                if off is None:
                    inline = off_payload.ljust(4, b'\0')
                else:
                    rel = off - extra_offset
                    extra_patch[rel:rel + len(off_payload)] = off_payload
            if off is None:
                fp.write(struct.pack('<HHI', tag, typ, count) + inline)
            else:
                fp.write(struct.pack('<HHII', tag, typ, count, off))
        fp.write(struct.pack('<I', 0))
        fp.write(bytes(extra_patch))
        fp.write(b'\0' * (data_offset - (extra_offset + len(extra))))
        fp.write(data.tobytes())


def _read_ifd(buf, bo):
    magic, = struct.unpack(bo + 'H', buf[2:4])
    if magic != 42:
        raise ValueError('not a classic TIFF file')
    ifd, = struct.unpack(bo + 'I', buf[4:8])
    n, = struct.unpack(bo + 'H', buf[ifd:ifd + 2])
    tags = {}
    for i in range(n):
        e = buf[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
        tag, typ, count = struct.unpack(bo + 'HHI', e[:8])
        size = _TYPE_SIZE.get(typ, 1) * count
        if size <= 4:
            raw = e[8:8 + size]
        else:
            off, = struct.unpack(bo + 'I', e[8:12])
            raw = buf[off:off + size]
        if typ == 2:
            val = bytes(raw).split(b'\0')[0].decode('ascii', 'replace')
        elif typ in _TYPE_FMT:
            val = list(struct.unpack(bo + '{}{}'.format(count, _TYPE_FMT[typ]), raw))
        else:
            val = bytes(raw)
        tags[tag] = val
    return tags


def _unpackbits(data):
    out = bytearray()
    i, n = 0, len(data)
    while i < n:
        c = data[i]
        i += 1
        if c < 128:
            out += data[i:i + c + 1]
            i += c + 1
        elif c > 128:
            out += data[i:i + 1] * (257 - c)
            i += 1
    return bytes(out)


def _unlzw(data):
    """TIFF LZW (MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, 'early change')."""
    out = bytearray()
    table, width, prev = None, 9, None
    bitbuf = nbits = pos = 0
    n = len(data)
    while True:
        while nbits < width:
            if pos >= n:
                return bytes(out)
            bitbuf = ((bitbuf << 8) | data[pos]) & 0xFFFFFF
            pos += 1
            nbits += 8
        code = (bitbuf >> (nbits - width)) & ((1 << width) - 1)
        nbits -= width
        if code == 257:
            return bytes(out)
        if code == 256:
            table = [bytes((i,)) for i in range(256)] + [b'', b'']
            width, prev = 9, None
            continue
        if table is None:
            raise ValueError('LZW stream does not start with a clear code')
        if prev is None:
            entry = table[code]
        elif code < len(table):
            entry = table[code]
            table.append(prev + entry[:1])
        else:
            entry = prev + prev[:1]
            table.append(entry)
        out += entry
        prev = entry
        if len(table) >= (1 << width) - 1 and width < 12:
            width += 1


def _decode_chunk(raw, compression, predictor, rows, cols, bo):
    """One strip or tile -> float32 (rows, cols)."""
    if compression == 1:
        data = bytes(raw)
    elif compression in (8, 32946):
        data = zlib.decompress(bytes(raw))
    elif compression == 5:
        data = _unlzw(bytes(raw))
    elif compression == 32773:
        data = _unpackbits(bytes(raw))
    else:
        raise ValueError('TIFF compression {} is not supported'.format(compression))
    if compression not in (5, 8, 32946):
        predictor = 1         # libtiff applies the Predictor tag only inside the LZW / Deflate codecs
    need = rows * cols * 4
    if len(data) < need:
        raise ValueError('strip/tile data is short: {} < {} bytes'.format(len(data), need))
    if predictor == 1:
        return np.frombuffer(data, dtype=bo + 'f4', count=rows * cols).reshape(rows, cols)
    if predictor == 2:        # horizontal differencing of the 32-bit words
        u = np.frombuffer(data, dtype=bo + 'u4', count=rows * cols).reshape(rows, cols)
        u = np.cumsum(u, axis=1, dtype=np.uint64).astype(np.uint32)
        return u.view(np.float32)
    if predictor == 3:        # floating-point predictor: byte planes (most significant first), differenced bytewise
        b = np.frombuffer(data, dtype=np.uint8, count=need).reshape(rows, cols * 4)
        b = np.cumsum(b, axis=1, dtype=np.uint64).astype(np.uint8)
        be = np.ascontiguousarray(b.reshape(rows, 4, cols).transpose(0, 2, 1))      # (rows, cols, 4) big-endian bytes
        return be.view('>f4').reshape(rows, cols).astype(np.float32)
    raise ValueError('TIFF predictor {} is not supported'.format(predictor))


def read_dsm_tif(file):
    assert (os.path.exists(file))
    with open(file, 'rb') as fp:
        buf = fp.read()
    bo = {b'II': '<', b'MM': '>'}.get(buf[:2])
    if bo is None:
        raise ValueError('not a TIFF file: {}'.format(file))
    tags = _read_ifd(buf, bo)
    width, height = tags[_T_WIDTH][0], tags[_T_LENGTH][0]
    assert (tags.get(_T_SPP, [1])[0] == 1)     # dsm is only one band (:59)
    assert (tags.get(_T_BITS, [0])[0] == 32 and tags.get(_T_SAMPLE_FORMAT, [1])[0] == 3)    # float32 (:62-63)
    compression = tags.get(_T_COMPRESSION, [1])[0]
    predictor = tags.get(_T_PREDICTOR, [1])[0]
    if tags.get(_T_PLANAR, [1])[0] != 1:
        raise ValueError('planar configuration 2 is not supported: {}'.format(file))
    image = np.zeros((height, width), dtype=np.float32)
    if _T_TILE_OFFSETS in tags:
        tw, tl = tags[_T_TILE_WIDTH][0], tags[_T_TILE_LENGTH][0]
        across = (width + tw - 1) // tw
        for i, (off, nbytes) in enumerate(zip(tags[_T_TILE_OFFSETS], tags[_T_TILE_BYTES])):
            tile = _decode_chunk(buf[off:off + nbytes], compression, predictor, tl, tw, bo)
            r0, c0 = (i // across) * tl, (i % across) * tw
            if r0 >= height:
                break
            image[r0:r0 + tl, c0:c0 + tw] = tile[:height - r0, :width - c0]       # edge tiles are padded
    else:
        rps = min(tags.get(_T_ROWS_PER_STRIP, [height])[0], height)
        for i, (off, nbytes) in enumerate(zip(tags[_T_STRIP_OFFSETS], tags[_T_STRIP_BYTES])):
            r0 = i * rps
            rows = min(rps, height - r0)
            if rows <= 0:
                break
            image[r0:r0 + rows] = _decode_chunk(buf[off:off + nbytes], compression, predictor, rows, width, bo)

    nodata = None
    if _T_GDAL_NODATA in tags:
        try:
            nodata = float(tags[_T_GDAL_NODATA])
        except ValueError:
            nodata = None
    # to ease later processing, replace nodata regions with nan (:69-72)
    if nodata is not None:
        mask = np.isclose(image, nodata)
        image[mask] = np.nan

    if _T_PIXEL_SCALE in tags and _T_TIEPOINT in tags:
        sx, sy = tags[_T_PIXEL_SCALE][0], tags[_T_PIXEL_SCALE][1]
        tp = tags[_T_TIEPOINT]
        geo = (tp[3] - tp[0] * sx, sx, 0.0, tp[4] + tp[1] * sy, 0.0, -sy)
    elif _T_TRANSFORM in tags:
        m = tags[_T_TRANSFORM]
        geo = (m[3], m[0], m[1], m[7], m[4], m[5])
    else:
        geo = (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    proj = ''
    keys = tags.get(_T_GEOKEYS)
    if keys:
        for i in range(4, len(keys) - 3, 4):
            if keys[i] == 3072 and keys[i + 1] == 0:
                epsg = keys[i + 3]
                if 32601 <= epsg <= 32660:
                    proj = _utm_wkt(epsg - 32600, 'N')
                elif 32701 <= epsg <= 32760:
                    proj = _utm_wkt(epsg - 32700, 'S')
    meta = {'AREA_OR_POINT': 'Area'} if keys and 1025 in keys[4::4] else {}
    # a TIFF without GeoTIFF keys has no projection string (GDAL would return ''; parse_proj_str needs a UTM name)
    zone_number, hemisphere = parse_proj_str(proj) if proj else (None, None)
    # return a meta dict (:76-99)
    meta_dict = {
        'geo': geo,
        'proj': proj,
        'meta': meta,
        'img_width': width,
        'img_height': height,
        'nodata': nodata,
        'zone_number': zone_number,
        'hemisphere': hemisphere,
        'ul_easting': geo[0],
        'ul_northing': geo[3],
        'east_resolution': geo[1],
        'north_resolution': abs(geo[5])
    }
    meta_dict['lr_easting'] = meta_dict['ul_easting'] + (meta_dict['img_width'] - 1) * meta_dict['east_resolution']
    meta_dict['lr_northing'] = meta_dict['ul_northing'] - (meta_dict['img_height'] - 1) * meta_dict['north_resolution']
    meta_dict['area_width'] = meta_dict['lr_easting'] - meta_dict['ul_easting']
    meta_dict['area_height'] = meta_dict['ul_northing'] - meta_dict['lr_northing']
    meta_dict['alt_min'] = float(np.nanmin(image))  # for json serialization
    meta_dict['alt_max'] = float(np.nanmax(image))
    return image, meta_dict
