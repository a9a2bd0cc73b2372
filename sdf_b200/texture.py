"""Texture-backed 2-D nodes: ``text`` / ``image`` / ``measure_text`` / ``measure_image`` (sdf/text.py:17-131) and the
``Mesh`` name (sdf/mesh.py:9).

``text`` and ``image`` rasterise a mask with Pillow, turn it into a signed Euclidean distance texture with scipy
(positive outside, negative inside, scaled to world units) and answer queries by bilinear interpolation of that
texture; outside the texture's rectangle the distance to the half-size rectangle stands in (sdf/text.py:100-118).
The evaluator has no opcode for a texture fetch, so these are HOST nodes of kind 'opaque', like a user-written
``@sdf2`` closure: ``extrude`` / ``revolve`` / CSG / rigid transforms around them are combined by
``core._compose_around_closure`` (closure-free operands still evaluate on the GPU) and the batch volumes are meshed
on the device (``core._generate_opaque``).  Pillow and scipy are imported on first use, as optional as they are for
the reference.

``Mesh`` keeps its container half (loading, bounding box, affine helpers); ``Mesh.sdf`` needs pyopenvdb (absent
here as it is for the reference) and raises.
"""
import numpy as np

from . import d2

PIXELS = 2 ** 22


def _open(thing):
    """sdf/text.py:11-16: a path, an ndarray, or anything ``np.array`` understands."""
    from PIL import Image
    if isinstance(thing, str):
        return Image.open(thing)
    if isinstance(thing, (np.ndarray, np.generic)):
        return Image.fromarray(thing)
    return Image.fromarray(np.array(thing))


def _fit(aspect, width, height):
    """World size from an aspect ratio and at most one given side; height 1 when neither is (sdf/text.py:22-28)."""
    if width is None and height is None:
        height = 1
    if width is None:
        width = height * aspect
    if height is None:
        height = width / aspect
    return width, height


def measure_text(name, text, width=None, height=None):
    from PIL import ImageFont
    l, t, r, b = ImageFont.truetype(name, 96).getbbox(text)
    return _fit((r - l) / (b - t), width, height)


def measure_image(thing, width=None, height=None):
    w, h = _open(thing).size
    return _fit(w / h, width, height)


class DistanceTexture:
    """The closure behind ``text`` / ``image``: a signed distance texture over a world rectangle.

    ``field``  (th, tw) float64, world-scaled signed distance per pixel
    ``margin`` (px, py) pixels of padding around the pw x ph window that maps to the world rectangle
    ``lo, hi`` world corners of that window; image row 0 is the TOP (v is flipped)
    """

    def __init__(self, mask, margin, width, height):
        import scipy.ndimage as nd
        inside = np.asarray(mask, dtype=bool)
        th, tw = inside.shape
        px, py = margin
        self.window = (tw - 2 * px, th - 2 * py)
        width, height = _fit(self.window[0] / self.window[1], width, height)
        field = np.zeros(inside.shape)
        field[inside] = -nd.distance_transform_edt(inside)[inside]
        field[~inside] = nd.distance_transform_edt(~inside)[~inside]
        field *= width / tw
        self.field = field
        self.margin = (px, py)
        self.size = (width, height)
        self.lo = (-width / 2, -height / 2)
        self.hi = (width / 2, height / 2)

    def pixel_coords(self, p):
        (x0, y0), (x1, y1) = self.lo, self.hi
        u = (p[:, 0] - x0) / (x1 - x0)
        v = 1 - (p[:, 1] - y0) / (y1 - y0)
        return u * self.window[0] + self.margin[0], v * self.window[1] + self.margin[1]

    def fetch(self, i, j):
        """Bilinear interpolation of ``field`` at column i, row j (sdf/text.py:120-143: the neighbours are clamped
        to the texture, the weights are taken against the CLAMPED indices)."""
        a = self.field
        c0 = np.floor(i).astype(int)
        r0 = np.floor(j).astype(int)
        c1 = np.clip(c0 + 1, 0, a.shape[1] - 1)
        r1 = np.clip(r0 + 1, 0, a.shape[0] - 1)
        c0 = np.clip(c0, 0, a.shape[1] - 1)
        r0 = np.clip(r0, 0, a.shape[0] - 1)
        return ((c1 - i) * (r1 - j) * a[r0, c0] + (c1 - i) * (j - r0) * a[r1, c0]
                + (i - c0) * (r1 - j) * a[r0, c1] + (i - c0) * (j - r0) * a[r1, c1])

    def fallback(self, p):
        """d2.rectangle((width / 2, height / 2)) at the origin (sdf/d2.py:107-120), used off the texture."""
        q = np.abs(p[:, :2]) - np.array(self.size) / 4
        return np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(np.maximum(q[:, 0], q[:, 1]), 0)

    def __call__(self, p):
        p = np.asarray(p, dtype=np.float64)
        i, j = self.pixel_coords(p)
        d = self.fetch(i, j)
        th, tw = self.field.shape
        off = (i < 0) | (i >= tw - 1) | (j < 0) | (j >= th - 1)
        d[off] = self.fallback(p)[off]
        return d


def _texture(im, margin, width, height, pixels):
    """Mask image -> closure (sdf/text.py:66-118): shrink to at most ``pixels`` texels, threshold, transform."""
    tw, th = im.size
    px, py = margin
    factor = (pixels / (tw * th)) ** 0.5
    if factor < 1:
        tw, th = int(round(tw * factor)), int(round(th * factor))
        px, py = int(round(px * factor)), int(round(py * factor))
        im = im.resize((tw, th))
    return DistanceTexture(np.array(im.convert('1')), (px, py), width, height)


@d2.sdf2
def text(font_name, text, width=None, height=None, pixels=PIXELS, points=512):
    from PIL import Image, ImageDraw, ImageFont
    font = ImageFont.truetype(font_name, points)
    l, t, r, b = font.getbbox(text)
    px, py = int((r - l) * 0.2), int((b - t) * 0.2)                 # 20 % padding on every side
    im = Image.new('L', (r - l + 1 + 2 * px, b - t + 1 + 2 * py))
    ImageDraw.Draw(im).text((px - l, py - t), text, font=font, fill=255)
    return _texture(im, (px, py), width, height, pixels)


@d2.sdf2
def image(thing, width=None, height=None, pixels=PIXELS):
    return _texture(_open(thing).convert('L'), (0, 0), width, height, pixels)


class Mesh:
    """sdf/mesh.py:9: a triangle mesh as an SDF source.  The container half -- loading, bounding box, the affine
    helpers (sdf/mesh.py:10-64) -- works; ``sdf()`` builds its level set with pyopenvdb in the reference
    (sdf/mesh.py:66-112), which exists neither here nor in the reference's own environment, and raises."""

    def __init__(self, points, triangles):
        self.points = np.asarray(points)
        self.triangles = np.asarray(triangles)

    @classmethod
    def from_file(cls, path):
        try:
            import meshio
        except ImportError:
            from . import mesh_io
            m = mesh_io.read(path)
            return cls(m.points, m.faces)
        m = meshio.read(path)
        return cls(m.points, m.cells[0].data)

    @property
    def bounding_box(self):
        return tuple(self.points.min(axis=0).tolist()), tuple(self.points.max(axis=0).tolist())

    @property
    def size(self):
        lo, hi = self.bounding_box
        return tuple((np.array(hi) - np.array(lo)).tolist())

    def transformed(self, matrix):
        """Homogeneous 4 x 4 (or 3 x 4) ``matrix`` applied to every vertex."""
        m = np.array(matrix, dtype=np.float64)
        return Mesh(self.points @ m[:3, :3].T + m[:3, 3], self.triangles)

    def scaled(self, scale):
        s = np.empty(3)
        s[...] = scale
        return self.transformed(np.hstack([np.diag(s), np.zeros((3, 1))]))

    def translated(self, offset):
        return self.transformed(np.hstack([np.eye(3), np.reshape(np.asarray(offset, np.float64), (3, 1))]))

    def positioned(self, position, anchor):
        lo, hi = (np.array(v) for v in self.bounding_box)
        return self.translated(np.asarray(position) - (lo + (hi - lo) * np.asarray(anchor)))

    def centered(self):
        return self.positioned((0, 0, 0), (0.5, 0.5, 0.5))

    def sdf(self, voxel_size, half_width=None):
        raise NotImplementedError(
            'Mesh.sdf() is not provided by sdf_b200: the reference builds the level set with pyopenvdb '
            '(sdf/mesh.py:66-82), which is outside the GPU mesh-generation path')
