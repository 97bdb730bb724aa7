"""-m gpu: the HTTP service twin (rabitq_service) against the reference's contract (crates/service/src/main.rs:36-88):
routes, JSON shapes, metrics text -- and that micro-batched answers equal per-request `query()` results."""
import http.client
import json
import os
import signal
import socket
import subprocess
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _wait_up(port, proc, timeout=60):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if proc.poll() is not None:
            raise RuntimeError("service exited: " + proc.stderr.read())
        try:
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=2)
            c.request("GET", "/health")
            r = c.getresponse()
            if r.status == 200 and r.read() == b"Ok":
                return
        except OSError:
            time.sleep(0.1)
    raise RuntimeError("service did not come up")


def test_service_contract_and_micro_batching(case_d128, tmp_path):
    import rabitq_b200 as rb
    from rabitq_b200 import build as bld

    case = case_d128
    d = tmp_path / "saved"
    case["oracle"].dump_to_dir(str(d))
    port = _free_port()
    proc = subprocess.Popen([bld.SERVICE, "-d", str(d), "-p", str(port), "-b", "bucket", "-k", "key", "-c", str(tmp_path / "cache"),
                             "--batch-wait-us", "2000"], stderr=subprocess.PIPE, text=True)
    try:
        _wait_up(port, proc)
        g = rb.RaBitQ.load_from_dir(str(d), device=0)
        q = case["queries"]
        direct = [g.query(q[i], 16, 10) for i in range(q.shape[0])]
        results = [None] * q.shape[0]

        def client(lo, hi):
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=30)
            for i in range(lo, hi):  # keep-alive: several requests on one connection
                body = json.dumps({"query": [float(x) for x in q[i]], "top_k": 10, "probe": 16})
                c.request("POST", "/query", body=body, headers={"content-type": "application/json"})
                r = c.getresponse()
                assert r.status == 200 and r.getheader("content-type") == "application/json"
                results[i] = json.loads(r.read())
            c.close()

        th = [threading.Thread(target=client, args=(i * 8, (i + 1) * 8)) for i in range(q.shape[0] // 8)]
        [t.start() for t in th]
        [t.join() for t in th]
        for i, r in enumerate(results):
            assert set(r.keys()) == {"ids", "scores"}                      # struct Response, main.rs:62-66
            assert r["ids"] == [x[1] for x in direct[i]]
            assert np.array_equal(np.array(r["scores"], np.float32).view(np.uint32),
                                  np.array([x[0] for x in direct[i]], np.float32).view(np.uint32))
        c = http.client.HTTPConnection("127.0.0.1", port, timeout=10)
        c.request("GET", "/")
        assert c.getresponse().read() == b"Ok"
        c.request("GET", "/metrics")
        m = c.getresponse().read().decode()
        assert m.startswith(f"query: {q.shape[0]}, rough: ") and "precise: " in m and m.endswith("cache miss: 0")   # metrics.rs:30-41
        ref = case["oracle"].query_batch(q, 16, 10)
        assert f"rough: {ref['rough']}, precise: {ref['precise']}," in m
        # errors: malformed body -> 422 (axum's Json rejection), wrong dimension -> 500 (the reference's handler panics), 404, 405
        c.request("POST", "/query", body="{\"query\": [1, 2], \"top_k\": 10}", headers={"content-type": "application/json"})
        r = c.getresponse(); r.read()
        assert r.status == 422
        c.request("POST", "/query", body=json.dumps({"query": [0.0] * 7, "top_k": 10, "probe": 4}))
        r = c.getresponse(); body = r.read().decode()
        assert r.status == 500 and "dim" in body
        # hostile parameters are rejected before any buffer is sized from them, and the service stays up
        for bad in ({"query": [0.0] * 128, "top_k": 4294967295, "probe": 4}, {"query": [0.0] * 128, "top_k": 0, "probe": 4},
                    {"query": [0.0] * 128, "top_k": 10, "probe": 0}, {"query": [], "top_k": 10, "probe": 4}):
            c.request("POST", "/query", body=json.dumps(bad))
            r = c.getresponse(); r.read()
            assert r.status == 422, bad
        c.request("POST", "/query", body='{"query": [0.0], "top_k": -1, "probe": 4}')
        r = c.getresponse(); r.read()
        assert r.status == 422
        c.request("POST", "/query", body='{"query": [0.0], "top_k": 99999999999999999999, "probe": 4}')
        r = c.getresponse(); r.read()
        assert r.status == 422
        c2 = http.client.HTTPConnection("127.0.0.1", port, timeout=10)   # oversized Content-Length: refused, connection closed
        c2.putrequest("POST", "/query"); c2.putheader("content-length", str(1 << 30)); c2.endheaders()
        r = c2.getresponse(); r.read()
        assert r.status == 413
        c2.close()
        c.request("GET", "/health")
        assert c.getresponse().read() == b"Ok"
        c.request("GET", "/nope")
        r = c.getresponse(); r.read()
        assert r.status == 404
        c.request("GET", "/query")
        r = c.getresponse(); r.read()
        assert r.status == 405
        c.close()
        g.close()
    finally:
        proc.send_signal(signal.SIGTERM)                                    # graceful shutdown, main.rs:17-30
        try:
            proc.wait(timeout=20)
        except subprocess.TimeoutExpired:
            proc.kill()
    err = proc.stderr.read()
    assert proc.returncode == 0, err
    assert "Server listening on 0.0.0.0:" in err and "Shutting down" in err
    # concurrent clients were really answered in batches
    line = [l for l in err.splitlines() if "requests in" in l][0]
    n_req, n_batches = int(line.split("answered ")[1].split()[0]), int(line.split(" in ")[1].split()[0])
    assert n_req == q.shape[0] + 1 and n_batches < n_req
