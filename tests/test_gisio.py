"""File forms of the reference's GIS inputs without GDAL / OGR (spinterps_b200/gisio.py):
ESRI shapefile polygons -> rings, ESRI ASCII grids and GeoTIFFs -> array rasters, and the path-taking
setters of SpInterpMain (interp/data.py:349-494, interp/main.py:291-343) on top of them."""
import struct

import numpy as np
import pytest

from spinterps_b200 import gisio


def _write_shp(path, shapes, shp_type=5):
    """shapes: list of (list of rings) or None (null shape); rings closed by the writer."""
    recs = b''
    for k, rings in enumerate(shapes):
        if rings is None:
            content = struct.pack('<i', 0)
        else:
            rings = [np.vstack([r, r[:1]]) if not np.array_equal(r[0], r[-1]) else r for r in rings]
            pts = np.vstack(rings)
            parts = np.cumsum([0] + [len(r) for r in rings[:-1]])
            content = struct.pack('<i', shp_type)
            content += struct.pack('<4d', pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(),
                                   pts[:, 1].max())
            content += struct.pack('<ii', len(rings), len(pts))
            content += np.asarray(parts, dtype='<i4').tobytes()
            content += np.ascontiguousarray(pts, dtype='<f8').tobytes()
            if shp_type == 15:          # PolygonZ: z range + z values (ignored by the reader)
                content += struct.pack('<2d', 0.0, 0.0) + np.zeros(len(pts), '<f8').tobytes()
        recs += struct.pack('>ii', k + 1, len(content) // 2) + content
    hdr = struct.pack('>i', 9994) + b'\x00' * 20 + struct.pack('>i', (100 + len(recs)) // 2)
    hdr += struct.pack('<ii', 1000, shp_type) + struct.pack('<8d', *([0.0] * 8))
    assert len(hdr) == 100
    path.write_bytes(hdr + recs)


SQ = np.array([[0.0, 0.0], [0.0, 10.0], [10.0, 10.0], [10.0, 0.0]])
HOLE = np.array([[4.0, 4.0], [6.0, 4.0], [6.0, 6.0], [4.0, 6.0]])
TRI = np.array([[20.5, 1.25], [30.0, 2.0], [25.0, 9.75]])


@pytest.mark.parametrize('shp_type', [5, 15])
def test_shapefile_rings(tmp_path, shp_type):
    p = tmp_path / 'polys.shp'
    _write_shp(p, [[SQ, HOLE], None, [TRI]], shp_type)
    rings = gisio.read_shp_polygons(p)
    # every ring becomes a polygon (misc.py:221-286), closed, in file order
    assert len(rings) == 3
    for got, exp in zip(rings, (SQ, HOLE, TRI)):
        assert np.array_equal(got, np.vstack([exp, exp[:1]]))
    bad = tmp_path / 'bad.shp'
    bad.write_bytes(b'\x00' * 120)
    with pytest.raises(ValueError):
        gisio.read_shp_polygons(bad)
    pts = tmp_path / 'points.shp'
    _write_shp(pts, [[SQ]], 5)
    raw = bytearray(pts.read_bytes())
    raw[32:36] = struct.pack('<i', 1)           # a point file
    pts.write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        gisio.read_shp_polygons(pts)


def test_ascii_grid(tmp_path):
    rng = np.random.default_rng(0)
    vals = np.round(rng.normal(500.0, 100.0, (5, 7)), 3)
    vals[1, 2] = -9999.0
    body = '\n'.join(' '.join(repr(float(v)) for v in row) for row in vals)
    a = tmp_path / 'dem.asc'
    a.write_text(f'ncols 7\nnrows 5\nxllcorner 1000.5\nyllcorner 2000.25\ncellsize 250\n'
                 f'NODATA_value -9999\n{body}\n')
    r = gisio.read_ascii_grid(a)
    assert np.array_equal(r['values'], vals) and r['values'].dtype == np.float64
    assert (r['x_min'], r['y_max'], r['cell_size'], r['ndv']) == (1000.5, 2000.25 + 5 * 250.0,
                                                                  250.0, -9999.0)
    b = tmp_path / 'dem_center.asc'
    b.write_text(f'NCOLS 7\nNROWS 5\nXLLCENTER 1125.5\nYLLCENTER 2125.25\nCELLSIZE 250\n{body}\n')
    r2 = gisio.read_raster(b)
    assert (r2['x_min'], r2['y_max'], r2['ndv']) == (1000.5, 3250.25, None)
    assert np.array_equal(r2['values'], vals)
    with pytest.raises(ImportError):
        gisio.read_raster(tmp_path / 'dem.img')


def _write_geotiff(path, arr, tags, compression=None):
    from PIL import Image, TiffImagePlugin
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    for tag, (typ, val) in tags.items():
        ifd[tag] = val
        ifd.tagtype[tag] = typ
    kw = {} if compression is None else {'compression': compression}
    Image.fromarray(arr).save(path, tiffinfo=ifd, **kw)


def test_geotiff(tmp_path):
    pytest.importorskip('PIL')
    rng = np.random.default_rng(1)
    elev = rng.normal(600.0, 150.0, (37, 53)).astype(np.float32)
    elev[3, 4] = -9999.0
    DOUBLE, ASCII, SHORT = 12, 2, 3
    base = {33550: (DOUBLE, (1000.0, 1000.0, 0.0)),
            33922: (DOUBLE, (0.0, 0.0, 0.0, 280000.5, 5650000.25, 0.0)),
            42113: (ASCII, '-9999')}
    for comp in (None, 'tiff_deflate', 'tiff_lzw'):
        p = tmp_path / f'elev_{comp}.tif'
        _write_geotiff(p, elev, base, comp)
        r = gisio.read_raster(p)
        assert r['values'].dtype == np.float64 and np.array_equal(r['values'], elev.astype(np.float64))
        assert (r['x_min'], r['y_max'], r['cell_size'], r['ndv']) == (280000.5, 5650000.25, 1000.0,
                                                                      -9999.0)
    # tiepoint at another pixel, PixelIsPoint (half a cell towards north-west), integer data
    dem = rng.integers(-50, 2500, (20, 31)).astype(np.int32)
    tags = {33550: (DOUBLE, (250.0, 250.0, 0.0)),
            33922: (DOUBLE, (2.0, 3.0, 0.0, 1000.0, 9000.0, 0.0)),
            34735: (SHORT, (1, 1, 0, 2, 1024, 0, 1, 1, 1025, 0, 1, 2))}
    p = tmp_path / 'dem_point.tif'
    _write_geotiff(p, dem, tags)
    r = gisio.read_geotiff(p)
    assert np.array_equal(r['values'], dem.astype(np.float64)) and r['ndv'] is None
    assert (r['x_min'], r['y_max'], r['cell_size']) == (1000.0 - 2 * 250.0 - 125.0,
                                                        9000.0 + 3 * 250.0 + 125.0, 250.0)
    # ModelTransformation instead of scale + tiepoint
    mat = (500.0, 0.0, 0.0, 12000.0, 0.0, -500.0, 0.0, 48000.0, 0.0, 0.0, 0.0, 0.0,
           0.0, 0.0, 0.0, 1.0)
    p = tmp_path / 'dem_matrix.tif'
    _write_geotiff(p, dem, {34264: (DOUBLE, mat)})
    r = gisio.read_geotiff(p)
    assert (r['x_min'], r['y_max'], r['cell_size']) == (12000.0, 48000.0, 500.0)
    p = tmp_path / 'plain.tif'
    _write_geotiff(p, dem, {})
    with pytest.raises(ValueError):
        gisio.read_geotiff(p)


def test_main_setters_take_the_files(tmp_path):
    from spinterps_b200 import SpInterpMain
    shp = tmp_path / 'catchments.shp'
    _write_shp(shp, [[SQ * 1000.0], [TRI * 1000.0]])
    asc = tmp_path / 'elev.asc'
    asc.write_text('ncols 4\nnrows 3\nxllcorner -5000\nyllcorner -6000\ncellsize 20000\n'
                   'NODATA_value -1\n1 2 3 4\n5 6 -1 8\n9 10 11 12\n')
    m = SpInterpMain(verbose=False)
    m.set_cell_selection_parameters(shp, 5000.0, True, 1000)
    assert m._cell_sel_prms_set and m._ipoly_flag and m._poly_shp == shp.absolute()
    assert (m._stn_bdist, m._cell_bdist) == (5000.0, 1000.0)
    assert len(m._poly_rings) == 2
    assert np.array_equal(m._poly_rings[1], np.vstack([TRI, TRI[:1]]) * 1000.0)
    with pytest.raises(NotImplementedError):
        m.set_cell_selection_parameters(shp, 5000.0, True, 1000, 0.1)
    with pytest.raises(AssertionError):
        m.set_cell_selection_parameters(tmp_path / 'missing.shp', 5000.0, True, 1000)
    m.turn_external_drift_kriging_on([asc])
    (ras,) = m._drft_rass
    assert m._edk_flag and ras['values'].shape == (3, 4) and ras['ndv'] == -1.0
    assert (ras['x_min'], ras['y_max'], ras['cell_size']) == (-5000.0, 54000.0, 20000.0)
    m.set_alignment_raster(asc)
    assert m._algn_ras_set_flag
    assert m._algn_ras == dict(x_min=-5000.0, y_max=54000.0, cell_size=20000.0, n_rows=3, n_cols=4)
