"""text() / image() (sdf_b200/texture.py, host nodes) and the host composition around them against vectors from the
REAL reference (tests/golden/make_texture_golden.py imports /root/reference/sdf with Pillow + scipy).

The closures themselves and every composition node are plain numpy on the host and must be BIT-identical to the
reference.  Closure-free operands (rounded_box, slab, circle ...) evaluate on the GPU in the product; here, without
one, the C-ABI evaluation is replaced by the numpy simulator of the register machine (tests/prog_sim.py), which
is float64 but not rounding-identical to the reference closures -- those trees compare to 1e-12.
"""
import os

import numpy as np
import pytest

import sdf_b200 as S
from sdf_b200 import _abi, core, ir
import prog_sim

PIL = pytest.importorskip('PIL')
pytest.importorskip('scipy.ndimage')

FONT = '/usr/local/cuda-12.9/compute-sanitizer/docs/_static/css/fonts/fontawesome-webfont.ttf'


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'texture.npz'))


@pytest.fixture()
def simulated_evaluator(monkeypatch):
    def eval_f64(self, xyz, out, n, device=0, stream=None):
        out[:] = prog_sim.run(self.program, xyz)
    monkeypatch.setattr(_abi.ProgramHandle, 'eval_f64', eval_f64)


def test_image_closure_bit_identical_to_reference(gold):
    blobs, bars, p2 = gold['mask_blobs'], gold['mask_bars'], gold['p2']
    cases = {
        'blobs_default': S.image(blobs),
        'blobs_width': S.image(blobs, width=1.2),
        'bars_height': S.image(bars, height=0.5),
        'bars_small': S.image(bars, width=1.5, pixels=3000),          # the down-scaling branch (text.py:70-74)
    }
    for key, f in cases.items():
        assert isinstance(f, S.d2.SDF2) and ir.node_of(f).kind == 'opaque'
        got = f(p2)
        assert got.shape == (len(p2), 1)
        assert np.array_equal(got.reshape(-1), gold['img2_' + key]), key
    # some of the points must be off the texture (rectangle fallback) and some on it
    tex = cases['blobs_default'].f
    i, j = tex.pixel_coords(p2)
    off = (i < 0) | (i >= blobs.shape[1] - 1) | (j < 0) | (j >= blobs.shape[0] - 1)
    assert 0.1 < off.mean() < 0.9


def test_measure_image_and_text(gold):
    bars, blobs = gold['mask_bars'], gold['mask_blobs']
    got = np.array([S.measure_image(bars), S.measure_image(bars, width=3), S.measure_image(blobs, height=2)]).reshape(-1)
    assert np.array_equal(got, gold['measure_image'])
    if 'measure_text' in gold.files and os.path.exists(FONT):
        got = np.array([S.measure_text(FONT, 'Hi'), S.measure_text(FONT, 'sdf', width=2)]).reshape(-1)
        assert np.array_equal(got, gold['measure_text'])


@pytest.mark.skipif(not os.path.exists(FONT), reason='the TrueType file the golden vectors were made with is not in this image')
def test_text_closure_bit_identical_to_reference(gold):
    p2 = gold['p2']
    assert np.array_equal(S.text(FONT, 'Hi')(p2).reshape(-1), gold['text2'])
    assert np.array_equal(S.text(FONT, 'sdf', width=1.5, pixels=20000, points=128)(p2).reshape(-1), gold['text2_small'])


def test_closure_only_trees_bit_identical_to_reference(gold):
    """Trees whose leaves are all closures need no evaluator: every node is the host composition."""
    blobs, p3 = gold['mask_blobs'], gold['p3']
    trees = {
        'extrude': S.image(blobs).extrude(0.3),
        'moved2d': S.image(blobs).rotate(0.4).translate((0.1, -0.2)).scale(0.8).extrude(0.25),
        'revolve': S.image(blobs, height=0.6).translate((0.5, 0)).revolve(),
    }
    for key, f in trees.items():
        assert isinstance(f, S.d3.SDF3)
        assert np.array_equal(f(p3).reshape(-1), gold['tree_' + key]), key


def test_mixed_trees_match_reference(gold, simulated_evaluator):
    """examples/image.py and examples/text.py shapes: library nodes (evaluator) combined with texture nodes (host)."""
    blobs, bars, p3 = gold['mask_blobs'], gold['mask_bars'], gold['p3']
    trees = {
        'example': S.rounded_box((0.8, 1.2, 0.1), 0.05) | (S.image(blobs).extrude(1) & S.slab(z0=0, z1=0.075)),
        'carved': S.box((1.2, 0.7, 0.3)) - S.image(bars, width=1.0).extrude(1).translate((0, 0, 0.2)),
        'extrude_to': S.image(blobs).extrude_to(S.d2.circle(0.3), 0.6, S.ease.in_out_quad),
        'union2d': (S.image(blobs) | S.d2.circle(0.2).translate((0.3, 0.3)).k(0.1)).extrude(0.2).orient(S.X),
    }
    if 'text_tree' in gold.files and os.path.exists(FONT):
        trees['text_tree'] = S.rounded_box((1.4, 1.0, 0.2), 0.05) - S.text(FONT, 'Hi', height=0.5).extrude(1)
    for key, f in trees.items():
        with pytest.raises(ir.NotLowerableError):
            ir.lower(f)
        got = f(p3).reshape(-1)
        ref = gold[key if key == 'text_tree' else 'tree_' + key]
        assert np.max(np.abs(got - ref)) <= 1e-12, (key, np.max(np.abs(got - ref)))


def test_nodes_that_cannot_sit_above_a_texture_say_so(gold):
    f = S.image(gold['mask_blobs']).extrude(0.3)
    with pytest.raises(ir.NotLowerableError, match='sits above a Python closure'):
        f.twist(1.0)(gold['p3'])


def test_mesh_container_follows_the_reference(tmp_path):
    """sdf/mesh.py:10-64 (the half of Mesh that needs no pyopenvdb): values from the formulas there, by hand."""
    from sdf_b200 import mesh_io
    pts = np.array([[0, 0, 0], [2, 0, 0], [0, 4, 0], [0, 0, 6]], np.float64)
    tris = np.array([[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]])
    m = S.Mesh(pts, tris)
    assert m.size == (2.0, 4.0, 6.0) and m.bounding_box == ((0.0, 0.0, 0.0), (2.0, 4.0, 6.0))
    assert m.centered().bounding_box == ((-1.0, -2.0, -3.0), (1.0, 2.0, 3.0))
    assert m.scaled(2).size == (4.0, 8.0, 12.0) and m.scaled((1, 0.5, 2)).size == (2.0, 2.0, 12.0)
    assert m.translated((1, 2, 3)).bounding_box[0] == (1.0, 2.0, 3.0)
    assert m.positioned((10, 10, 10), (1, 1, 1)).bounding_box[1] == (10.0, 10.0, 10.0)
    swap = [[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 5], [0, 0, 0, 1]]
    assert m.transformed(swap).bounding_box == ((0.0, 0.0, 5.0), (4.0, 2.0, 11.0))
    assert m.transformed(swap).triangles is m.triangles or np.array_equal(m.transformed(swap).triangles, tris)
    path = str(tmp_path / 't.ply')
    mesh_io.IndexedMesh(pts, [('triangle', tris)]).write(path)
    back = S.Mesh.from_file(path)
    assert np.array_equal(back.points, pts) and np.array_equal(back.triangles, tris)
    with pytest.raises(NotImplementedError, match='pyopenvdb'):
        m.sdf(voxel_size=0.25)


def test_skip_rule_and_bounds_run_on_texture_trees(gold):
    """core._estimate_bounds and the host skip test (core._skip_host) only call the tree: closure-only tree, no GPU."""
    f = S.image(gold['mask_blobs'], height=1.0).extrude(0.2)
    (x0, y0, z0), (x1, y1, z1) = core._estimate_bounds(f)
    assert -0.6 < x0 < -0.05 and 0.05 < x1 < 0.6 and -0.7 < y0 < -0.2 and 0.2 < y1 < 0.7
    assert -0.2 < z0 <= -0.1 and 0.1 <= z1 < 0.2
    far = (np.linspace(5, 5.2, 9), np.linspace(5, 5.2, 9), np.linspace(5, 5.2, 9))
    near = (np.linspace(-0.1, 0.1, 9),) * 3
    assert core._skip_host(f, far) and not core._skip_host(f, near)


def test_oracle_evaluates_texture_trees_bit_identically(gold):
    """The oracle (numpy restatement of every library node) calls a texture node like any other closure, so it can
    check the GPU route for these trees (tests/test_gpu_zzz_texture.py): pin it on the same vectors first."""
    from oracle import sdf_ref
    blobs, bars, p3 = gold['mask_blobs'], gold['mask_bars'], gold['p3']
    trees = {
        'example': S.rounded_box((0.8, 1.2, 0.1), 0.05) | (S.image(blobs).extrude(1) & S.slab(z0=0, z1=0.075)),
        'carved': S.box((1.2, 0.7, 0.3)) - S.image(bars, width=1.0).extrude(1).translate((0, 0, 0.2)),
        'revolve': S.image(blobs, height=0.6).translate((0.5, 0)).revolve(),
        'extrude_to': S.image(blobs).extrude_to(S.d2.circle(0.3), 0.6, S.ease.in_out_quad),
        'union2d': (S.image(blobs) | S.d2.circle(0.2).translate((0.3, 0.3)).k(0.1)).extrude(0.2).orient(S.X),
    }
    for key, f in trees.items():
        assert np.array_equal(sdf_ref.evaluate(f, p3).reshape(-1), gold['tree_' + key]), key


def test_generate_flow_on_a_texture_tree_with_device_stand_ins(gold, simulated_evaluator, monkeypatch):
    """generate() on an examples/image.py-shaped tree, with the two device calls of that route replaced by host
    stand-ins (evaluator -> register-machine simulator, sdfb200_mc_volume -> the oracle's marching cubes + the
    world mapping of core.py:58-60): batches, the skip rule, the one-plane tail batch and the counters must come
    out as the oracle's generate() has them.  The same tree runs on the real device in tests/test_gpu_zzz_texture.py."""
    from oracle import sdf_ref

    def mc_volume(self, volume, X, Y, Z, device=0, stream=None):
        try:
            pts = sdf_ref.marching_cubes_soup(np.asarray(volume, np.float64))
        except Exception:
            return np.zeros((0, 9), np.float32)
        pts = pts * np.array([X[1] - X[0], Y[1] - Y[0], Z[1] - Z[0]]) + np.array([X[0], Y[0], Z[0]])
        return pts.astype(np.float32).reshape(-1, 9)
    monkeypatch.setattr(_abi.ProgramHandle, 'mc_volume', mc_volume)
    m = gold['mask_blobs']
    w, h = S.measure_image(m)
    f = S.rounded_box((w * 1.1, h * 1.1, 0.1), 0.05) | (S.image(m).extrude(1) & S.slab(z0=0, z1=0.075))
    kw = dict(step=0.0127, bounds=((-0.41, -0.61, -0.11), (0.413, 0.6, 0.13)))
    # a batch larger than the device's 33^3 volume is meshed in pieces: same soup as the oracle's big batches
    big, ib = S.generate(f, verbose=False, return_info=True, batch_size=48, **kw)
    rbig, sb = sdf_ref.generate(f, batch_size=48, **kw)
    assert ib['batches'] == sb['batches'] == 4 and (ib['skipped'], ib['empty'], ib['nonempty']) == (sb['skipped'], sb['empty'], sb['nonempty'])
    from util_cmp import compare_soups
    rep = compare_soups(big, rbig)
    # (a piece is mapped to world space from its own first sample: vertices may differ in the last float32 bit)
    assert rep['count_equal'] and rep['max_vertex_dist'] <= 2e-7 and rep['exact_triangles'] >= 0.9 * rep['n_ref'], rep
    got, info = S.generate(f, verbose=False, return_info=True, **kw)
    ref, st = sdf_ref.generate(f, **kw)
    assert info['dims'] == (65, 96, 19) and info['batches'] == 9
    assert (info['skipped'], info['empty'], info['nonempty']) == (st['skipped'], st['empty'], st['nonempty']) == (0, 3, 6)
    assert len(got) == len(ref) == 3 * st['triangles'] and st['triangles'] > 20000
    from util_cmp import compare_soups
    rep = compare_soups(got, ref)
    assert rep['max_vertex_dist'] <= 1e-5 and rep['exact_triangles'] >= 0.999 * rep['n_ref'], rep
