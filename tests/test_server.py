"""The optional FastAPI route: validation on CPU (no compute), a full request on the GPU."""
import base64
import json
import os

import numpy as np
import pytest

from conftest import ROOT

fastapi = pytest.importorskip("fastapi")
from fastapi import FastAPI  # noqa: E402
from fastapi.testclient import TestClient  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
DAT = GOLD["parser"]["naca0012_selig_test_main"]["text"].encode()


@pytest.fixture()
def client(monkeypatch):
    from aerolab_lbm import dat, server
    # stand-alone deployment: plain reader instead of the host application's parser
    monkeypatch.setattr(server, "resolve_parser", lambda p: dat.read_plain_dat)
    monkeypatch.setattr(server, "_limiter", server._RateLimiter(10 ** 6, 60.0))   # see test_rate_limit
    app = FastAPI()
    app.include_router(server.router)
    return TestClient(app)


def post(client, data=None, name="n0012.dat", content=DAT):
    return client.post("/lbm/run/", files={"file": (name, content, "application/octet-stream")}, data=data or {})


@pytest.mark.parametrize("data,frag", [
    ({"alpha": "45"}, "Alpha must be"), ({"u0": "0.5"}, "U0 must be"), ({"tau": "0.4"}, "tau must be"),
    ({"nx": "4"}, "Lattice must be"), ({"steps": "0"}, "steps must be"), ({"field": "foo"}, "field must be"),
])
def test_validation_400(client, data, frag):
    r = post(client, data)
    assert r.status_code == 400 and frag in r.json()["detail"]


def test_rejects_non_dat_and_oversize(client):
    assert post(client, name="foil.txt").status_code == 400
    assert post(client, content=b"1 0\n" * 400000).status_code == 400


def test_rate_limit_like_upload_airfoil(client, monkeypatch):
    """main.py:543-545 limits /upload_airfoil/ to 5 requests per minute and client; so does this route."""
    from aerolab_lbm import server
    monkeypatch.setattr(server, "_limiter", server._RateLimiter(server.RATE_LIMIT, server.RATE_WINDOW_S))
    codes = [post(client, {"alpha": "45"}).status_code for _ in range(server.RATE_LIMIT + 2)]
    assert codes == [400] * server.RATE_LIMIT + [429, 429]
    lim = server._RateLimiter(2, 10.0)
    assert [lim.check("a", now=t) for t in (0.0, 1.0, 2.0, 10.5, 11.5, 12.0)] == [True, True, False, True, True, False]
    assert lim.check("b", now=2.0)


def test_too_few_points_and_work_cap(client):
    r = post(client, content=b"1 0\n0.5 0.05\n0 0\n0.5 -0.05\n1 0\n")
    assert r.status_code == 400 and ("Too few points" in r.json()["detail"] or "fewer than 10" in r.json()["detail"])
    r = post(client, {"nx": "8192", "ny": "8192", "steps": "200000"})
    assert r.status_code == 400 and "cell updates" in r.json()["detail"]


def test_nan_becomes_null():
    from aerolab_lbm import server
    import json as _json
    out = server._json_safe({"a": float("nan"), "b": [1.0, float("inf")], "c": {"d": np.float64("nan")}, "e": 2})
    assert out == {"a": None, "b": [1.0, None], "c": {"d": None}, "e": 2}
    _json.dumps(out, allow_nan=False)


def test_health_route(client):
    r = client.get("/lbm/health")
    assert r.status_code == 200 and "cuda_devices" in r.json()


def test_no_gpu_means_500_not_a_cpu_answer(client):
    import aerolab_lbm as al
    if al.device_count() > 0:
        pytest.skip("a GPU is present")
    r = post(client, {"steps": "4"})
    assert r.status_code == 500 and "no CPU fallback" in r.json()["detail"]


@pytest.mark.gpu
def test_full_request_on_gpu(client, built_lib):
    from oracle import geometry as ogeo
    from oracle import lbm as olbm
    r = post(client, {"alpha": "5", "steps": "120", "return_field": "true", "field": "cp"})
    assert r.status_code == 200, r.text
    body = r.json()
    assert body["success"] and body["num_points"] == 13 and body["steps"] == 120
    o = olbm.OracleTunnel(320, 160)
    o.apply_geometry(ogeo.round_coords(GOLD["parser"]["naca0012_selig_test_main"]["coords"]), 5.0)
    for k in range(1, 31):
        o.step(4)
        if k % 3 == 0:
            o.compute_forces()
    assert body["coefficients"]["CL"] == pytest.approx(o.cl_smooth, rel=1e-12)
    assert body["coefficients"]["CD"] == pytest.approx(o.cd_smooth, rel=1e-12)
    assert body["separation"]["state"] == o.stall_state()[0]
    assert body["reynolds"] == pytest.approx(o.reynolds())
    rgba = np.frombuffer(base64.b64decode(body["field"]["data"]), np.uint8).reshape(body["field"]["shape"])
    o.update_fields()
    assert rgba.shape == (160, 320, 4)
    assert np.array_equal(rgba, olbm.rgba(o.mask, o.render(1), 1))
