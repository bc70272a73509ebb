"""CPU suite for the host-only parts of s4g_release_b200/file_logger.py (reference utils/file_logger_cls.py:121-170): the
"jet" colour map restated from matplotlib's published segment table, the ASCII PLY writers, the frame glyphs.  The dump
itself (softmax, collision check on the device): tests/test_file_logger_gpu.py."""
import numpy as np

from s4g_release_b200.file_logger import frame_glyphs, jet_colors, write_ply_mesh, write_ply_points


def test_jet_colour_map_anchor_values():
    """matplotlib's `jet`: dark blue at 0, cyan / yellow around the middle, dark red at 1; every channel in [0, 1]"""
    c = jet_colors(np.array([0.0, 0.125, 0.375, 0.5, 0.64, 0.9, 1.0]), levels=1024)
    assert c.shape == (7, 3) and (c >= 0).all() and (c <= 1).all()
    assert np.allclose(c[0], [0.0, 0.0, 0.5], atol=5e-3)          # jet(0)
    assert np.allclose(c[-1], [0.5, 0.0, 0.0], atol=5e-3)         # jet(1)
    assert np.allclose(c[1], [0.0, 0.0, 1.0], atol=2e-2)          # blue plateau begins (b reaches 1 at 0.11, g leaves 0 at 0.125)
    assert c[2][1] > 0.98 and c[2][2] > 0.85                      # cyan side: green saturated at 0.375, blue still high
    assert abs(c[3][0] - c[3][2]) < 0.05 and c[3][1] == 1.0       # the middle is green with equal red / blue
    assert c[4][0] > 0.9 and c[4][1] > 0.98 and c[4][2] < 0.05    # yellow around 0.64-0.66
    assert c[5][0] > 0.95 and c[5][1] < 0.05                      # red end: green gone by 0.91
    # out-of-range values clamp to the end colours like a matplotlib colour map does
    assert np.allclose(jet_colors(np.array([-1.0, 2.0])), jet_colors(np.array([0.0, 1.0])))


def test_jet_is_piecewise_linear_between_break_points():
    x = np.linspace(0.38, 0.63, 50)   # red ramps 0 -> 1 between 0.35 and 0.66, green is 1 (0.375 .. 0.64), blue ramps 1 -> 0
    c = jet_colors(x, levels=4096)
    assert (np.diff(c[:, 0]) >= -1e-9).all() and (np.diff(c[:, 2]) <= 1e-9).all() and np.allclose(c[:, 1], 1.0)
    slope = np.polyfit(x, c[:, 0], 1)[0]
    assert abs(slope - 1.0 / (0.66 - 0.35)) < 0.05


def test_ply_writers_round_trip(tmp_path):
    pts = np.array([[0.1, 0.2, 0.3], [1.0, -2.0, 3.5]])
    col = np.array([[0.0, 0.5, 1.0], [1.0, 1.0, 0.0]])
    path = tmp_path / "p.ply"
    write_ply_points(str(path), pts, col)
    lines = path.read_text().splitlines()
    assert lines[0] == "ply" and lines[1] == "format ascii 1.0" and "element vertex 2" in lines
    body = lines[lines.index("end_header") + 1:]
    assert len(body) == 2
    v = [float(t) for t in body[1].split()]
    assert np.allclose(v[:3], pts[1]) and v[3:] == [255, 255, 0]
    mesh = tmp_path / "m.ply"
    write_ply_mesh(str(mesh), np.vstack([pts, [[0, 0, 0]]]), np.vstack([col, [[0, 0, 0]]]), np.array([[0, 1, 2]]))
    lines = mesh.read_text().splitlines()
    assert "element face 1" in lines and lines[-1] == "3 0 1 2"


def test_frame_glyphs_layout():
    """12 vertices per drawn point (point, midpoint, origin, 3 x (origin, axis tip, axis tip offset)), one face = the y-axis
    triangle; every `stride`-th point only (file_logger_cls.py:121-158)"""
    rs = np.random.RandomState(0)
    pts, t = rs.rand(7, 3), rs.rand(7, 3)
    R = np.stack([np.linalg.qr(rs.randn(3, 3))[0] for _ in range(7)])
    verts, col, tri = frame_glyphs(pts, t, R, stride=2)
    n = 4  # points 0, 2, 4, 6
    assert verts.shape == (12 * n, 3) and col.shape == (12 * n, 3) and tri.shape == (n, 3)
    assert np.allclose(verts[0], pts[0]) and np.allclose(verts[2], t[0]) and np.allclose(verts[12], pts[2])
    assert np.allclose(verts[7], t[0] + R[0][:, 1] * 0.01 + R[0][:, 2] * 0.001)   # y-axis marker
    assert (tri[1] == 12 + np.array([6, 7, 8])).all()
    assert (col[0] == [1, 0, 0]).all() and (col[3] == [0, 1, 0]).all() and (col[6] == [1, 1, 0]).all() and (col[9] == [0, 0, 1]).all()
