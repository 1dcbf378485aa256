"""Robustness of the parsers that read bytes from the network or from users: the cluster / stage JSON reader (config server PUT, runner
`update` / `lost` control messages), peer lists, host lists and hostfiles. Whatever the input, the answer is a value or a Python
exception — never a crash — and what was written can be read back."""
import json

import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from kungfu_b200 import _kfrt

ports = st.integers(min_value=1, max_value=65535)
ips = st.tuples(st.integers(1, 254), st.integers(0, 255), st.integers(0, 255), st.integers(1, 254)).map(lambda t: ".".join(map(str, t)))
peers = st.tuples(ips, ports).map(lambda t: f"{t[0]}:{t[1]}")


def _ip_to_int(ip):
    a, b, c, d = (int(x) for x in ip.split("."))
    return (a << 24) | (b << 16) | (c << 8) | d


@settings(max_examples=300, deadline=None)
@given(st.text(max_size=200))
def test_cluster_json_never_crashes_on_text(s):
    for f in (_kfrt.cluster_from_json, _kfrt.stage_from_json, _kfrt.cluster_validate):
        try:
            f(s)
        except Exception:  # noqa: BLE001 - any Python exception is fine
            pass


json_values = st.recursive(st.none() | st.booleans() | st.integers(-2**40, 2**40) | st.floats(allow_nan=False, allow_infinity=False) | st.text(max_size=8),
                           lambda c: st.lists(c, max_size=4) | st.dictionaries(st.sampled_from(["Runners", "Workers", "IPv4", "Port", "Version", "Cluster", "x"]), c, max_size=4),
                           max_leaves=12)


@settings(max_examples=300, deadline=None)
@given(json_values)
def test_cluster_json_never_crashes_on_wellformed_json_of_the_wrong_shape(v):
    s = json.dumps(v)
    for f in (_kfrt.cluster_from_json, _kfrt.stage_from_json):
        try:
            f(s)
        except Exception:  # noqa: BLE001
            pass


@settings(max_examples=200, deadline=None)
@given(st.lists(peers, min_size=1, max_size=6, unique=True), st.lists(peers, min_size=0, max_size=12, unique=True))
def test_cluster_json_round_trip(runners, workers):
    doc = {"Runners": [{"IPv4": _ip_to_int(p.split(":")[0]), "Port": int(p.split(":")[1])} for p in runners],
           "Workers": [{"IPv4": _ip_to_int(p.split(":")[0]), "Port": int(p.split(":")[1])} for p in workers]}
    c = _kfrt.cluster_from_json(json.dumps(doc))
    assert list(c["runners"]) == runners and list(c["workers"]) == workers


@settings(max_examples=300, deadline=None)
@given(st.text(alphabet=st.characters(blacklist_categories=("Cs",)), max_size=120))
def test_peer_and_host_lists_never_crash(s):
    for f in (_kfrt.parse_peer_list, _kfrt.parse_hostfile):
        try:
            f(s)
        except Exception:  # noqa: BLE001
            pass


@pytest.mark.parametrize("bad", ['{"Runners": 3, "Workers": []}', '{"Runners": [{"IPv4": "x", "Port": 1}], "Workers": []}', '[1,2', '{"Runners": [], "Workers": [{"IPv4": 1}]}',
                                 '{"Version": "one", "Cluster": {}}', "\x00\x01\x02", "{" * 2000, '{"Runners": [[[[[[[[[[]]]]]]]]]], "Workers": null}'])
def test_malformed_cluster_documents_are_rejected_with_an_exception_or_an_empty_cluster(bad):
    try:
        c = _kfrt.cluster_from_json(bad)
    except Exception:  # noqa: BLE001
        return
    assert isinstance(c, dict)


@pytest.mark.parametrize("opener", ["[", "{\"a\":", "[{\"Workers\":"])
def test_deep_nesting_is_an_error_not_a_stack_overflow(opener):
    """100 000 nested containers used to overflow the C++ stack of whoever parsed them (a runner reading an `update`, a config server
    reading a PUT); the reader now refuses documents deeper than 64 levels."""
    with pytest.raises(Exception, match="nesting too deep|json"):
        _kfrt.cluster_from_json(opener * 100_000)
    with pytest.raises(Exception):
        _kfrt.stage_from_json(opener * 1_000_000)


def test_config_server_shrugs_off_hostile_http():
    """The config server listens where anybody can reach it: endless headers, absurd Content-Length, binary junk and deeply nested JSON bodies
    are dropped or answered with 400; afterwards it still serves."""
    import socket

    srv = _kfrt.ConfigServer("/config")
    port = srv.start("127.0.0.1", 0, "")
    url = f"http://127.0.0.1:{port}/config"
    try:
        good = _kfrt.cluster_json("127.0.0.1:38080", "127.0.0.1:10000,127.0.0.1:10001")
        assert _kfrt.http_request("PUT", url, good)[0] == 200

        def raw(payload: bytes, read: bool = True) -> bytes:
            s = socket.create_connection(("127.0.0.1", port), timeout=5)
            try:
                try:
                    s.sendall(payload)
                except OSError:
                    pass  # the server may hang up while we are still talking: that is the point
                if not read:
                    return b""
                s.settimeout(8)
                try:
                    return s.recv(4096)
                except OSError:
                    return b""
            finally:
                s.close()

        raw(b"GET /config HTTP/1.1\r\nX: " + b"a" * (200 << 10))  # 200 KiB of header without an end: dropped, not buffered for ever
        raw(b"PUT /config HTTP/1.1\r\nContent-Length: 99999999999999\r\n\r\n{}")
        raw(b"PUT /config HTTP/1.1\r\nContent-Length: -5\r\n\r\n{}")
        raw(bytes(range(256)) * 16)
        deep = b"[" * 200_000
        ans = raw(b"PUT /config HTTP/1.1\r\nContent-Length: %d\r\n\r\n" % len(deep) + deep)
        assert ans.startswith(b"HTTP/1.1 400"), ans[:80]
        st, body = _kfrt.http_request("GET", url)
        assert st == 200 and json.loads(body) == json.loads(good) and srv.version() == 1
    finally:
        srv.stop()
