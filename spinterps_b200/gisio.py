"""Readers for three simple GIS file formats, so that the path-taking setters of the
reference (``set_cell_selection_parameters(polygons_shapefile, ...)``, interp/data.py:349-461;
``turn_external_drift_kriging_on([raster paths])``, interp/main.py:291-343;
``set_alignment_raster(path)``, interp/data.py:463-494) work without GDAL / OGR for

* ESRI shapefiles (``.shp``) holding polygons -- every ring of every shape becomes one
  polygon, which is what the reference makes of them: ``misc.linearize_sub_polys``
  (misc.py:221-286) recurses into the rings of a multi-ring geometry and wraps each one
  ("Polygons with holes do not get interpolated! Just accept them anyway.");
* ESRI ASCII grids (``.asc``): ``ncols / nrows / xllcorner|xllcenter / yllcorner|yllcenter /
  cellsize / NODATA_value`` followed by the rows from north to south;
* single-band GeoTIFFs (``.tif``) when Pillow is importable (pixels by libtiff, the
  georeferencing tags are interpreted here).

Anything else (projections, rotated rasters, other vector formats) still needs GDAL: the
array forms of the setters take data read by whatever the caller has.
Format: ESRI Shapefile Technical Description (1998), main file: a 100-byte header (file code
9994 big-endian, version 1000 and shape type little-endian), then records of (record number,
content length in 16-bit words; both big-endian) + content (shape type, box, numParts,
numPoints, parts[], points[] -- little-endian).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_POLYGON_TYPES = (5, 15, 25)        # Polygon, PolygonZ, PolygonM


def read_shp_polygons(path):
    """Rings of every polygon record of a shapefile: list of (n, 2) float64 arrays (closed:
    first vertex repeated), in file order."""
    raw = Path(path).read_bytes()
    if len(raw) < 100 or struct.unpack('>i', raw[:4])[0] != 9994:
        raise ValueError(f'{path}: not an ESRI shapefile')
    n_words = struct.unpack('>i', raw[24:28])[0]
    if struct.unpack('<i', raw[28:32])[0] != 1000:
        raise ValueError(f'{path}: unsupported shapefile version')
    shp_type = struct.unpack('<i', raw[32:36])[0]
    if shp_type not in _POLYGON_TYPES:
        raise ValueError(f'{path}: shape type {shp_type} is not a polygon type')
    end = min(len(raw), 2 * n_words)
    rings = []
    pos = 100
    while pos + 8 <= end:
        _, c_words = struct.unpack('>ii', raw[pos:pos + 8])
        beg, pos = pos + 8, pos + 8 + 2 * c_words
        if pos > end:
            raise ValueError(f'{path}: truncated record')
        rec_type = struct.unpack('<i', raw[beg:beg + 4])[0]
        if rec_type == 0:                       # null shape
            continue
        if rec_type not in _POLYGON_TYPES:
            raise ValueError(f'{path}: record of shape type {rec_type} in a polygon file')
        n_parts, n_pts = struct.unpack('<ii', raw[beg + 36:beg + 44])
        parts = np.frombuffer(raw, dtype='<i4', count=n_parts, offset=beg + 44)
        pts = np.frombuffer(raw, dtype='<f8', count=2 * n_pts,
                            offset=beg + 44 + 4 * n_parts).reshape(n_pts, 2)
        bounds = list(parts) + [n_pts]
        for a, b in zip(bounds[:-1], bounds[1:]):
            if not (0 <= a < b <= n_pts):
                raise ValueError(f'{path}: bad part index')
            rings.append(np.array(pts[a:b], dtype=np.float64))
    if not rings:
        raise ValueError(f'{path}: no polygons')
    return rings


def read_ascii_grid(path):
    """ESRI ASCII grid -> dict(values [rows, cols] float64 (row 0 = north), x_min, y_max,
    cell_size, ndv) -- the array form ``turn_external_drift_kriging_on`` and
    ``set_alignment_raster`` take."""
    hdr = {}
    with open(path, 'r') as fh:
        pos = fh.tell()
        while True:
            line = fh.readline()
            tok = line.split()
            if len(tok) == 2 and tok[0][0].isalpha():
                hdr[tok[0].lower()] = tok[1]
                pos = fh.tell()
                continue
            break
        fh.seek(pos)
        vals = np.loadtxt(fh, dtype=np.float64, ndmin=2)
    try:
        n_cols, n_rows = int(hdr['ncols']), int(hdr['nrows'])
        cs = float(hdr['cellsize'])
    except KeyError as exc:
        raise ValueError(f'{path}: not an ESRI ASCII grid (missing {exc})') from None
    if 'xllcorner' in hdr:
        x_min = float(hdr['xllcorner'])
    else:
        x_min = float(hdr['xllcenter']) - 0.5 * cs
    if 'yllcorner' in hdr:
        y_min = float(hdr['yllcorner'])
    else:
        y_min = float(hdr['yllcenter']) - 0.5 * cs
    vals = vals.reshape(-1)
    if vals.size != n_rows * n_cols:
        raise ValueError(f'{path}: {vals.size} values for {n_rows} x {n_cols} cells')
    ndv = float(hdr['nodata_value']) if 'nodata_value' in hdr else None
    return dict(values=vals.reshape(n_rows, n_cols), x_min=x_min, y_max=y_min + n_rows * cs,
                cell_size=cs, ndv=ndv)


def read_geotiff(path):
    """Single-band GeoTIFF -> dict(values, x_min, y_max, cell_size, ndv), what
    ``misc.get_ras_props`` + ``ReadAsArray`` give the reference (misc.py:630-686,
    interp/drift.py:55-83).  The pixels are decoded by Pillow (libtiff; strips or tiles, raw /
    LZW / deflate, 8 / 16 / 32-bit integers and 32-bit floats); the georeferencing is taken
    from the GeoTIFF tags here: ModelPixelScale (33550) + ModelTiepoint (33922), or an
    unrotated ModelTransformation (34264); a PixelIsPoint raster (GTRasterTypeGeoKey 1025 = 2
    in the GeoKeyDirectory 34735) is shifted by half a cell like GDAL's geotransform; the
    no-data value comes from GDAL_NODATA (42113)."""
    try:
        from PIL import Image
    except Exception as exc:  # noqa: BLE001
        raise ImportError('reading a GeoTIFF needs Pillow or GDAL; pass the raster as a '
                          'dict(values, x_min, y_max, cell_size, ndv)') from exc
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as img:
        tags = dict(img.tag_v2) if hasattr(img, 'tag_v2') else {}
        if len(img.getbands()) != 1:
            raise ValueError(f'{path}: {len(img.getbands())} bands, expected a single-band raster')
        vals = np.array(img)
    if vals.ndim != 2:
        raise ValueError(f'{path}: not a 2-D raster')
    scale = tags.get(33550)
    tie = tags.get(33922)
    mat = tags.get(34264)
    if scale is not None and tie is not None:
        sx, sy = float(scale[0]), float(scale[1])
        i, j, _, x, y = (float(v) for v in tie[:5])
        x_min, y_max = x - i * sx, y + j * sy
    elif mat is not None and len(mat) == 16:
        if float(mat[1]) != 0.0 or float(mat[4]) != 0.0:
            raise ValueError(f'{path}: rotated rasters are not supported')
        sx, sy = float(mat[0]), -float(mat[5])
        x_min, y_max = float(mat[3]), float(mat[7])
    else:
        raise ValueError(f'{path}: no GeoTIFF georeferencing tags')
    keys = tags.get(34735)
    if keys is not None:
        keys = [int(k) for k in keys]
        for q in range(4, len(keys) - 3, 4):
            if keys[q] == 1025 and keys[q + 1] == 0 and keys[q + 3] == 2:    # PixelIsPoint
                x_min -= 0.5 * sx
                y_max += 0.5 * sy
    if not np.isclose(sx, sy):
        raise ValueError(f'{path}: cells are not square ({sx}, {sy})')
    ndv = tags.get(42113)
    if ndv is not None:
        ndv = float(str(ndv).strip().strip('\x00'))
        if np.isnan(ndv):
            ndv = None
    return dict(values=np.ascontiguousarray(vals, dtype=np.float64), x_min=x_min, y_max=y_max,
                cell_size=sx, ndv=ndv)


def read_raster(path):
    """Dispatch on the file type: ESRI ASCII grids (no dependency) and GeoTIFFs (Pillow);
    everything else needs GDAL."""
    p = Path(path)
    if p.suffix.lower() in ('.asc', '.txt'):
        return read_ascii_grid(p)
    if p.suffix.lower() in ('.tif', '.tiff'):
        return read_geotiff(p)
    raise ImportError(f'reading {p.suffix or p.name} rasters needs GDAL; pass the raster as a '
                      'dict(values, x_min, y_max, cell_size, ndv), a GeoTIFF or an ESRI ASCII grid')
