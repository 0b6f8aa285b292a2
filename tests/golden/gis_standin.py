"""In-memory stand-ins for the two GIS libraries the reference's grid preparation calls
(osgeo.ogr / osgeo.gdal are absent from the build container).  TEST INFRASTRUCTURE of
``make_golden_prep.py`` only: they let the reference's OWN preparation code
(interp/prepare.py, interp/bdpolys.py, interp/drift.py, misc.py:221-540) run unmodified on
polygons and rasters held in arrays.  Only the calls that code makes are provided.

What the stand-ins decide themselves (and the real libraries would decide with GEOS /
GDAL): ``Geometry.Contains`` (even-odd crossing rule) and ``Geometry.Buffer`` (exact
distance instead of arcs of 30 segments per quadrant).  ``make_golden_prep.py`` only keeps
inputs on which those two cannot differ from GEOS: every tested point is farther from
every ring and from every buffer outline than the sagitta of GEOS's arcs.
"""
import numpy as np

wkbLineString = 2
wkbPolygon = 3
wkbMultiPolygon = 6
wkbLinearRing = 101


class Geometry:
    def __init__(self, gtype):
        self._t = gtype
        self._pts = []          # ring
        self._rings = []        # polygon
        self._buf = 0.0

    # construction -------------------------------------------------------
    def AddPoint(self, x, y, z=0.0):
        self._pts.append((float(x), float(y)))

    def AddGeometry(self, ring):
        self._rings.append(ring.Clone())

    def Clone(self):
        g = Geometry(self._t)
        g._pts = list(self._pts)
        g._rings = [r.Clone() for r in self._rings]
        g._buf = self._buf
        return g

    # queries ------------------------------------------------------------
    def GetGeometryType(self):
        return wkbLineString if self._t == wkbLinearRing else self._t

    def GetGeometryName(self):
        return {wkbLinearRing: 'LINEARRING', wkbPolygon: 'POLYGON'}.get(self._t, 'POINT')

    def GetGeometryCount(self):
        return len(self._rings)

    def GetGeometryRef(self, i):
        return self._rings[i]

    def GetPoints(self):
        return list(self._pts)

    def GetPointCount(self):
        return len(self._pts)

    def GetX(self):
        return self._pts[0][0]

    def GetY(self):
        return self._pts[0][1]

    def _xy(self):
        return np.asarray(self._rings[0]._pts, dtype=np.float64)

    def Area(self):
        r = self._xy()
        return 0.5 * abs(np.dot(r[:, 0], np.roll(r[:, 1], -1))
                         - np.dot(r[:, 1], np.roll(r[:, 0], -1)))

    def GetEnvelope(self):
        r = self._xy()
        b = self._buf
        return (r[:, 0].min() - b, r[:, 0].max() + b, r[:, 1].min() - b, r[:, 1].max() + b)

    def GetLinearGeometry(self):
        return self

    def SimplifyPreserveTopology(self, tol):
        raise NotImplementedError('stand-in: simplify_tolerance_ratio must stay 0')

    def Buffer(self, dist):
        g = self.Clone()
        g._buf += float(dist)
        return g

    def edge_distance(self, x, y):
        r = self._xy()
        a, b = r, np.roll(r, -1, axis=0)
        d = b - a
        l2 = (d * d).sum(axis=1)
        w = np.array([x, y]) - a
        t = np.where(l2 > 0, (w * d).sum(axis=1) / np.where(l2 > 0, l2, 1.0), 0.0)
        t = np.clip(t, 0.0, 1.0)
        q = w - t[:, None] * d
        return float(np.sqrt((q * q).sum(axis=1)).min())

    def Contains(self, pt):
        x, y = pt.GetX(), pt.GetY()
        r = self._xy()
        ax, ay = r[:, 0], r[:, 1]
        bx, by = np.roll(ax, -1), np.roll(ay, -1)
        strad = (ay > y) != (by > y)
        with np.errstate(divide='ignore', invalid='ignore'):
            xi = (bx - ax) * (y - ay) / (by - ay) + ax
        inside = bool(np.count_nonzero(strad & (x < xi)) & 1)
        if inside or self._buf <= 0.0:
            return inside
        return self.edge_distance(x, y) < self._buf


def CreateGeometryFromWkt(wkt):
    assert wkt.startswith('POINT (') and wkt.endswith(')'), wkt
    x, y = wkt[7:-1].split()
    g = Geometry(1)
    g.AddPoint(float(x), float(y))
    return g


def polygon_from_ring(xy):
    """Closed outer ring (first vertex repeated, as OGR hands shapefile rings out)."""
    xy = np.asarray(xy, dtype=np.float64)
    if not np.array_equal(xy[0], xy[-1]):
        xy = np.vstack([xy, xy[:1]])
    ring = Geometry(wkbLinearRing)
    for x, y in xy:
        ring.AddPoint(x, y)
    poly = Geometry(wkbPolygon)
    poly.AddGeometry(ring)
    return poly


class _Feature:
    def __init__(self, geom):
        self._g = geom

    def GetGeometryRef(self):
        return self._g


class _Layer:
    def __init__(self, polys):
        self._polys = polys

    def __iter__(self):
        return iter([_Feature(p) for p in self._polys])

    def GetExtent(self):
        env = np.array([p.GetEnvelope() for p in self._polys])
        return (env[:, 0].min(), env[:, 1].max(), env[:, 2].min(), env[:, 3].max())


class _VectorDS:
    def __init__(self, polys):
        self._lyr = _Layer(polys)

    def GetLayerCount(self):
        return 1

    def GetLayer(self, i):
        assert i == 0
        return self._lyr

    def Destroy(self):
        pass


class _Band:
    DataType = 7        # GDT_Float64

    def __init__(self, arr, ndv):
        self._a, self._ndv = arr, ndv

    def ReadAsArray(self):
        return self._a.copy()

    def GetNoDataValue(self):
        return self._ndv


class _RasterDS:
    RasterCount = 1

    def __init__(self, arr, x_min, y_max, cell, ndv):
        self._b = _Band(np.asarray(arr), ndv)
        self._gt = (float(x_min), float(cell), 0.0, float(y_max), 0.0, -float(cell))
        self.RasterYSize, self.RasterXSize = arr.shape

    def GetGeoTransform(self):
        return self._gt

    def GetRasterBand(self, i):
        assert i == 1
        return self._b

    def GetProjectionRef(self):
        return ''


class Registry:
    """``ogr`` and ``gdal`` in one object: ``Open(path)`` looks the path up in the
    vector / raster tables filled by the generating script."""
    Geometry = Geometry
    wkbLinearRing = wkbLinearRing
    wkbPolygon = wkbPolygon
    CE_None = 0
    CreateGeometryFromWkt = staticmethod(CreateGeometryFromWkt)

    def __init__(self):
        self.vectors = {}
        self.rasters = {}

    def UseExceptions(self):
        pass

    def add_polygons(self, path, rings):
        self.vectors[str(path)] = [polygon_from_ring(r) for r in rings]

    def add_raster(self, path, arr, x_min, y_max, cell, ndv):
        self.rasters[str(path)] = (np.asarray(arr), x_min, y_max, cell, ndv)

    def Open(self, path, *a):
        path = str(path)
        if path in self.vectors:
            return _VectorDS(self.vectors[path])
        if path in self.rasters:
            return _RasterDS(*self.rasters[path])
        return None
