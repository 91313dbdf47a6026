"""The TCP service speaks the reference's wire format (ralm/retriever/serialization_utils.py:17-292,
ralm/server/faiss_server.py:241-277).  CPU-only: a stand-in index checks the transport, the GPU test puts a real index
behind it."""
import os
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))


class _FakeIndex:
    """Deterministic stand-in: id = 1000 * row + rank (+ first list id), distance = rank + sum of the query."""
    d = 24

    def __init__(self):
        self.nprobe = 1

    def search(self, x, k):
        n = x.shape[0]
        I = (np.arange(n)[:, None] * 1000 + np.arange(k)[None, :]).astype(np.int64)
        D = (x.sum(1, keepdims=True) + np.arange(k)[None, :]).astype(np.float32)
        return D, I

    def search_preassigned(self, x, k, list_ids):
        D, I = self.search(x, k)
        return D, I + list_ids[:, :1]


def test_message_lengths_match_the_reference_formulas():
    from b200ivfpq import server as w
    assert w.request_message_length(32, 512) == 4 + 32 * 512 * 4                  # serialization_utils.py:17-18
    assert w.request_message_length_with_lists(32, 512, 16) == 16 + 32 * (512 * 4 + 16 * 8)   # :20-22
    assert w.answer_message_len(10, 32) == 32 * 10 * 12                           # :34-35
    q = np.arange(6, dtype=np.float32).reshape(2, 3)
    msg = w.encode_request(q, 7)
    assert msg[:4] == b"\x00\x00\x00\x07" and len(msg) == w.request_message_length(2, 3)     # big-endian k
    k, q2 = w.decode_request(msg, 2, 3)
    assert k == 7 and np.array_equal(q2, q)
    l = np.array([[5, -1], [2, 9]], np.int64)
    msg = w.encode_request_with_lists(q, l, 4)
    assert msg[:16] == b"".join(int(v).to_bytes(4, "big") for v in (2, 3, 2, 4))
    k, q2, l2 = w.decode_request_with_lists(msg, 2, 3, 2)
    assert k == 4 and np.array_equal(q2, q) and np.array_equal(l2, l)
    I = np.array([[3, 1], [-1, 8]], np.int64)
    D = np.array([[0.5, 1.5], [2.5, 3.5]], np.float32)
    ans = w.encode_answer(I, D)
    assert ans[:I.nbytes] == I.tobytes() and ans[I.nbytes:] == D.tobytes()        # ids first, then distances (:253-255)
    I2, D2 = w.decode_answer(ans, 2, 2)
    assert np.array_equal(I2, I) and np.array_equal(D2, D)
    with pytest.raises(ValueError):
        w.decode_request(msg, 2, 3)


@pytest.mark.parametrize("with_lists", [False, True])
def test_server_round_trips(with_lists):
    from b200ivfpq.server import B200Client, B200Server
    index = _FakeIndex()
    srv = B200Server(index, port=0, batch_size=4, dim=index.d, default_k=5, nprobe=3, request_with_lists=with_lists)
    t = threading.Thread(target=srv.start, daemon=True)
    t.start()
    cli = B200Client(srv.address[0], srv.address[1], batch_size=4, dim=index.d, k=5, nprobe=3)
    rng = np.random.default_rng(0)
    for it in range(3):
        q = rng.random((4, index.d), dtype=np.float32)
        lists = rng.integers(0, 50, size=(4, 3)).astype(np.int64) if with_lists else None
        out = cli.retrieve(q, lists)
        D, I = (index.search_preassigned(q, 5, lists) if with_lists else index.search(q, 5))
        assert np.array_equal(out["id"], I) and np.array_equal(out["dist"], D)
        assert out["id"].dtype == np.int64 and out["dist"].dtype == np.float32
    cli.close()
    t.join(timeout=5)
    assert not t.is_alive() and srv.served == 3 and index.nprobe == 3
    srv.close()


def test_server_survives_a_bad_request():
    """A request with another k is rejected BEFORE any search runs, answered with an empty result of the agreed length,
    and the loop keeps serving (the reference's server does not die on a bad request)."""
    import threading
    from b200ivfpq.server import B200Client, B200Server
    index = _FakeIndex()
    calls = []
    orig = index.search
    index.search = lambda x, k: (calls.append(k), orig(x, k))[1]
    srv = B200Server(index, port=0, batch_size=4, dim=index.d, default_k=5, nprobe=3)
    port = srv.server.getsockname()[1]
    t = threading.Thread(target=srv.start, kwargs={"max_requests": 2}, daemon=True)
    t.start()
    bad = B200Client("127.0.0.1", port, 4, index.d, k=7)
    bad.k = 7
    q = np.random.default_rng(0).random((4, index.d), dtype=np.float32)
    from b200ivfpq import server as w
    bad.sock.sendall(w.encode_request(q, 7)[:srv.query_msg_len])          # k = 7 in the header, same message length
    ans = w._recv_exact(bad.sock, w.answer_message_len(5, 4))
    I, D = w.decode_answer(ans, 5, 4)
    assert (I == -1).all() and calls == [], "a rejected request must not reach the index"
    bad.k = 5
    out = bad.retrieve(q)
    assert out["id"].shape == (4, 5) and calls == [5] and srv.rejected == 1
    bad.close()
    t.join(timeout=5)
    srv.close()


@pytest.mark.gpu
def test_server_with_a_real_index(oracle):
    import _util
    import b200ivfpq as faiss
    from b200ivfpq.server import B200Client, B200Server
    a = _util.make_index_arrays(oracle, 3, 64, 16, 16, 6000)
    index = faiss.IndexIVFPQ(faiss.IndexFlatL2(64), 64, 16, 16, 8)
    index.set_codebooks(a["coarse"], a["pq"])
    index.set_lists(a["offsets"], a["codes"], a["ids"])
    srv = B200Server(index, port=0, batch_size=8, default_k=10, nprobe=4)
    t = threading.Thread(target=srv.start, daemon=True)
    t.start()
    cli = B200Client(srv.address[0], srv.address[1], batch_size=8, dim=64, k=10)
    xq = _util.make_queries(5, a, 8)
    out = cli.retrieve(xq)
    cli.close()
    t.join(timeout=10)
    srv.close()
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 4, 10)
    _util.assert_bit_equal(out["dist"], Dr, "D over the wire")
    _util.assert_bit_equal(out["id"], Ir, "I over the wire")


def test_wire_bytes_equal_the_references_encoders():
    """tests/golden/wire_format.npz was produced by RUNNING the reference's serialization_utils.py
    (tests/golden/make_wire_golden.py): our encoders must emit exactly its bytes and our decoders must read its bytes
    back to the inputs."""
    from b200ivfpq import server as w
    g = np.load(os.path.join(ROOT, "tests", "golden", "wire_format.npz"))
    for tag in ("a", "b", "c"):
        batch, dim, nprobe, k = (int(v) for v in g[f"{tag}_shape"])
        q, lists, I, D = g[f"{tag}_q"], g[f"{tag}_lists"], g[f"{tag}_I"], g[f"{tag}_D"]
        req, reql, ans = g[f"{tag}_req"].tobytes(), g[f"{tag}_req_lists"].tobytes(), g[f"{tag}_ans"].tobytes()
        assert bytes(w.encode_request(q, k)) == req, tag
        assert bytes(w.encode_request_with_lists(q, lists, k)) == reql, tag
        assert bytes(w.encode_answer(I, D)) == ans, tag
        k2, q2 = w.decode_request(req, batch, dim)
        assert k2 == k and np.array_equal(q2, q)
        k3, q3, l3 = w.decode_request_with_lists(reql, batch, dim, nprobe)
        assert k3 == k and np.array_equal(q3, q) and np.array_equal(l3, lists)
        I2, D2 = w.decode_answer(ans, k, batch)
        assert np.array_equal(I2, I) and np.array_equal(D2, D)
        assert w.request_message_length(batch, dim) == len(req)
        assert w.request_message_length_with_lists(batch, dim, nprobe) == len(reql)
        assert w.answer_message_len(k, batch) == len(ans)


REF_RALM = "/root/reference/Chameleon/llm_inference_gpu"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_RALM, "ralm", "retriever", "retriever.py")),
                    reason="the reference is only mounted in the build container")
@pytest.mark.parametrize("with_lists", [False, True])
def test_the_references_own_client_talks_to_our_server(with_lists):
    """Drop-in check with CODE OF THE REFERENCE on the other end of the socket: its `ExternalRetriever`
    (ralm/retriever/retriever.py:68-183, imported from where it lies) connects to B200Server, sends its own encoded
    requests -- blocking `retrieve`, the with-lists variant, and the tik-tok pattern retrieve_send / poll /
    retrieve_recv (ralm_tiktok.py:129-192) -- and decodes our answers."""
    import time
    from b200ivfpq.server import B200Server
    sys.path.insert(0, REF_RALM)
    try:
        from ralm.retriever.retriever import ExternalRetriever
    finally:
        sys.path.remove(REF_RALM)
    index = _FakeIndex()
    batch, k, nprobe = 4, 5, 3
    srv = B200Server(index, port=0, batch_size=batch, dim=index.d, default_k=k, nprobe=nprobe,
                     request_with_lists=with_lists)
    t = threading.Thread(target=srv.start, daemon=True)
    t.start()
    cli = ExternalRetriever(host="127.0.0.1", port=srv.address[1], batch_size=batch, dim=index.d, default_k=k)
    rng = np.random.default_rng(1)
    for it in range(3):
        q = rng.random((batch, index.d), dtype=np.float32)
        lists = rng.integers(0, 50, size=(batch, nprobe)).astype(np.int64)
        if with_lists:
            D, I = index.search_preassigned(q, k, lists)
            if it < 2:
                ids, dists = cli.retrieve_with_lists(q, lists)
            else:
                cli.retrieve_with_lists_send(q, lists, k)
                ids, dists = cli.retrieve_recv(k)
        else:
            D, I = index.search(q, k)
            if it < 2:
                ids, dists = cli.retrieve(q, k)
            else:                                   # the asynchronous pattern the decoder loop uses
                cli.retrieve_send(q, k)
                deadline = time.time() + 10
                while not cli.poll() and time.time() < deadline:
                    time.sleep(0.001)
                assert cli.poll(), "answer never became readable"
                ids, dists = cli.retrieve_recv(k)
        assert np.array_equal(np.asarray(ids, np.int64).reshape(batch, k), I)
        assert np.array_equal(np.asarray(dists, np.float32).reshape(batch, k), D)
    cli.socket.close()
    t.join(timeout=5)
    assert not t.is_alive() and srv.served == 3
    srv.close()
