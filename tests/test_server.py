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
