"""CPU: host-side logic of the Python package (no compute calls)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, assert_bitwise
from oracle import geometry as ogeo

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def test_product_shapes_equal_oracle_bitwise():
    import aerolab_lbm as al
    assert set(al.SHAPES) == set(ogeo.SHAPES)
    for key in al.SHAPES:
        assert_bitwise(np.array(al.SHAPES[key]()), np.array(ogeo.SHAPES[key]()), key)
    assert_bitwise(np.array(al.naca_digits("NACA 6409")), np.array(ogeo.naca4(6, 4, 9, 50)), "digits")
    with pytest.raises(ValueError):
        al.naca_digits("23012")


def test_round_coords_matches_bridge():
    import aerolab_lbm as al
    pts = [[0.1234567891, -0.00000049], [1, 0]]
    assert al.round_coords(pts) == [[0.123457, -0.0], [1.0, 0.0]] == ogeo.round_coords(pts)


def test_slab_rows_partition():
    from aerolab_lbm.distributed import slab_rows
    for ny, world in ((16384, 8), (160, 3), (10, 4), (7, 7)):
        rows = [slab_rows(ny, world, r) for r in range(world)]
        assert rows[0][0] == 0 and sum(n for _, n in rows) == ny
        for (y0, n), (y1, _) in zip(rows, rows[1:]):
            assert y0 + n == y1
        assert max(n for _, n in rows) - min(n for _, n in rows) <= 1


def test_plain_dat_reader_agrees_with_reference_parser_on_clean_selig(tmp_path):
    """The committed fixture was produced by the REAL main.parse_dat_file (make_golden.py)."""
    from aerolab_lbm.dat import read_plain_dat
    fx = GOLD["parser"]["naca0012_selig_test_main"]
    p = tmp_path / "n0012.dat"
    p.write_text(fx["text"])
    coords, fixes = read_plain_dat(str(p))
    assert coords == fx["coords"] and len(coords) == 13
    with pytest.raises(ValueError):
        q = tmp_path / "short.dat"
        q.write_text("1 0\n0 0\n")
        read_plain_dat(str(q))


def test_reference_parser_bridge_when_app_present(tmp_path):
    """Beside the real application the tunnel uses main.parse_dat_file unmodified."""
    if not os.path.exists("/root/reference/main.py"):
        pytest.skip("reference application not present on this machine")
    from aerolab_lbm.dat import load_reference_parser
    parse = load_reference_parser("/root/reference")
    for fx in GOLD["parser"].values():
        p = tmp_path / "a.dat"
        p.write_text(fx["text"])
        coords, fixes = parse(str(p))
        assert [[float(x), float(y)] for x, y in coords] == fx["coords"] and list(fixes) == fx["fixes"]


def test_resolve_parser_error_message(monkeypatch):
    import sys
    from aerolab_lbm import dat
    monkeypatch.delenv("AEROLAB_APP_DIR", raising=False)
    monkeypatch.setattr(sys, "path", [p for p in sys.path if "reference" not in p])
    sys.modules.pop("main", None)
    with pytest.raises(ImportError) as e:
        dat.resolve_parser(None)
    assert "parse_dat_file" in str(e.value)
    assert dat.resolve_parser(dat.read_plain_dat) is dat.read_plain_dat


def test_resplit_rows_balances_and_conserves():
    """Static slab balancing (DistributedTunnel.rebalance): pure host arithmetic."""
    from aerolab_lbm.distributed import resplit_rows
    rows = [2048] * 8
    times = [1.0, 1.0, 1.0, 1.15, 1.15, 1.0, 1.0, 1.0]      # the slabs with the body are slower
    new = resplit_rows(rows, times, 16384)
    assert sum(new) == 16384 and min(new) >= 8
    assert new[3] < 2048 and new[4] < 2048 and new[0] > 2048
    # equal times: nothing changes
    assert resplit_rows(rows, [2.0] * 8, 16384) == rows
    # a fixed point of the iteration: time proportional to rows
    r2 = resplit_rows(new, [t * n / 2048 for t, n in zip(times, new)], 16384)
    assert sum(r2) == 16384 and all(abs(a - b) < 120 for a, b in zip(r2, new))
    # tiny lattices keep at least min_rows per slab
    assert min(resplit_rows([10, 10, 10], [1.0, 50.0, 1.0], 30, min_rows=4)) >= 4
