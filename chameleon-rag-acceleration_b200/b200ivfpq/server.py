"""TCP retrieval service with the reference's wire format.

The multi-process RALM deployment talks to its retriever over a socket: one connection, many fixed-size requests
(llm_inference_gpu/ralm/server/faiss_server.py:241-277, ralm/retriever/serialization_utils.py:17-292).  This module
keeps that wire format, so the reference's client (ralm/retriever/client.py / ExternalRetriever) can point at a B200
box unchanged:

    request               k (4 B, big endian) | queries f32[batch, dim]                        (C order, native endian)
    request with lists    batch, dim, nprobe, k (4 x 4 B, big endian) | queries f32[batch, dim] | list ids i64[batch, nprobe]
    answer                ids i64[batch, k] | distances f32[batch, k]

The server is a thin loop around index.search / search_preassigned; the search itself is the CUDA path (for the small
batches a RALM decode step sends, the five-launch CUDA-graph latency path).
"""
from __future__ import annotations

import socket
from typing import Optional, Tuple

import numpy as np

_BE = "big"


def request_message_length(batch_size: int, dim: int) -> int:
    return 4 + batch_size * dim * 4


def request_message_length_with_lists(batch_size: int, dim: int, nprobe: int) -> int:
    return 16 + batch_size * (dim * 4 + nprobe * 8)


def answer_message_len(k: int, batch_size: int) -> int:
    return batch_size * k * (8 + 4)


def encode_request(queries: np.ndarray, k: int) -> bytes:
    q = np.ascontiguousarray(queries, np.float32)
    return int(k).to_bytes(4, _BE) + q.tobytes(order="C")


def decode_request(msg: bytes, batch_size: int, dim: int) -> Tuple[int, np.ndarray]:
    if len(msg) != request_message_length(batch_size, dim):
        raise ValueError(f"request is {len(msg)} bytes, expected {request_message_length(batch_size, dim)}")
    k = int.from_bytes(msg[:4], _BE)
    return k, np.frombuffer(msg, dtype=np.float32, offset=4).reshape(batch_size, dim)


def encode_request_with_lists(queries: np.ndarray, list_ids: np.ndarray, k: int) -> bytes:
    q = np.ascontiguousarray(queries, np.float32)
    l = np.ascontiguousarray(list_ids, np.int64)
    if q.ndim != 2 or l.ndim != 2 or l.shape[0] != q.shape[0]:
        raise ValueError("queries (batch, dim) and list_ids (batch, nprobe) expected")
    head = b"".join(int(v).to_bytes(4, _BE) for v in (q.shape[0], q.shape[1], l.shape[1], k))
    return head + q.tobytes(order="C") + l.tobytes(order="C")


def decode_request_with_lists(msg: bytes, batch_size: int, dim: int, nprobe: int):
    if len(msg) != request_message_length_with_lists(batch_size, dim, nprobe):
        raise ValueError("request length does not match (batch, dim, nprobe)")
    b, d, p, k = (int.from_bytes(msg[4 * i:4 * i + 4], _BE) for i in range(4))
    if (b, d, p) != (batch_size, dim, nprobe):
        raise ValueError(f"request header ({b}, {d}, {p}) does not match the service's ({batch_size}, {dim}, {nprobe})")
    qbytes = batch_size * dim * 4
    queries = np.frombuffer(msg, dtype=np.float32, count=batch_size * dim, offset=16).reshape(batch_size, dim)
    lists = np.frombuffer(msg, dtype=np.int64, count=batch_size * nprobe, offset=16 + qbytes).reshape(batch_size, nprobe)
    return k, queries, lists


def encode_answer(indices: np.ndarray, distances: np.ndarray) -> bytes:
    i = np.ascontiguousarray(indices, np.int64)
    d = np.ascontiguousarray(distances, np.float32)
    if i.shape != d.shape or i.ndim != 2:
        raise ValueError("indices and distances must both be (batch, k)")
    return i.tobytes(order="C") + d.tobytes(order="C")


def decode_answer(msg: bytes, k: int, batch_size: int) -> Tuple[np.ndarray, np.ndarray]:
    if len(msg) != answer_message_len(k, batch_size):
        raise ValueError("answer length does not match (batch, k)")
    n = batch_size * k
    return (np.frombuffer(msg, dtype=np.int64, count=n).reshape(batch_size, k),
            np.frombuffer(msg, dtype=np.float32, count=n, offset=8 * n).reshape(batch_size, k))


def _recv_exact(sock: socket.socket, n: int) -> Optional[bytes]:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            return None if not buf else bytes(buf)
        buf += chunk
    return bytes(buf)


class B200Server:
    """FaissServer's protocol (faiss_server.py:170-277): fixed batch size, dimension, k (and nprobe) per service; one
    client connection at a time, many requests per connection."""

    def __init__(self, index, host: str = "127.0.0.1", port: int = 9091, batch_size: int = 32, dim: Optional[int] = None,
                 default_k: int = 10, nprobe: int = 32, request_with_lists: bool = False):
        self.index = index
        self.batch_size, self.default_k, self.nprobe = int(batch_size), int(default_k), int(nprobe)
        self.dim = int(dim if dim is not None else index.d)
        self.request_with_lists = bool(request_with_lists)
        self.index.nprobe = self.nprobe
        self.query_msg_len = (request_message_length_with_lists(self.batch_size, self.dim, self.nprobe)
                              if self.request_with_lists else request_message_length(self.batch_size, self.dim))
        self.server = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.server.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.server.bind((host, port))
        self.server.listen(1)
        self.address = self.server.getsockname()
        self.served = 0
        self.rejected = 0

    def retrieve(self, query: np.ndarray, k: Optional[int] = None):
        D, I = self.index.search(np.ascontiguousarray(query, np.float32), k or self.default_k)
        return {"id": np.asarray(I), "dist": np.asarray(D)}

    def retrieve_with_lists(self, query: np.ndarray, list_ids: np.ndarray, k: Optional[int] = None):
        D, I = self.index.search_preassigned(np.ascontiguousarray(query, np.float32), k or self.default_k,
                                             np.ascontiguousarray(list_ids, np.int64))
        return {"id": np.asarray(I), "dist": np.asarray(D)}

    def _empty_answer(self) -> bytes:
        """What a search that found nothing returns (ids -1, distances FLT_MAX), in the fixed answer length the client
        waits for: the reply to a request this service cannot serve."""
        ids = np.full((self.batch_size, self.default_k), -1, np.int64)
        dist = np.full((self.batch_size, self.default_k), np.finfo(np.float32).max, np.float32)
        return encode_answer(ids, dist)

    def handle(self, msg: bytes) -> bytes:
        """One request -> one answer.  The request's k is checked BEFORE anything is searched; a request the service
        cannot serve (wrong k, malformed lists, a failing search) is logged and answered with an empty result, and the
        loop keeps serving -- the reference's server does not die on a bad request either (faiss_server.py:241-277)."""
        try:
            if self.request_with_lists:
                k, queries, lists = decode_request_with_lists(msg, self.batch_size, self.dim, self.nprobe)
            else:
                k, queries = decode_request(msg, self.batch_size, self.dim)
                lists = None
            if k != self.default_k:
                raise ValueError(f"request asks for k = {k}, the service was started with k = {self.default_k}")
            out = self.retrieve_with_lists(queries, lists, k) if lists is not None else self.retrieve(queries, k)
            return encode_answer(out["id"], out["dist"])
        except Exception as e:      # noqa: BLE001 -- a service loop: report and keep going
            self.rejected += 1
            print(f"[B200Server] request rejected: {type(e).__name__}: {e}", flush=True)
            return self._empty_answer()

    def start(self, max_requests: Optional[int] = None):
        """Accept one connection and serve it until the client closes it (or max_requests have been answered)."""
        conn, _ = self.server.accept()
        conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        try:
            while max_requests is None or self.served < max_requests:
                msg = _recv_exact(conn, self.query_msg_len)
                if msg is None:
                    break
                if len(msg) != self.query_msg_len:
                    raise ConnectionError("client closed the connection in the middle of a request")
                conn.sendall(self.handle(msg))
                self.served += 1
        finally:
            conn.close()

    def close(self):
        self.server.close()


class B200Client:
    """The client side of the same protocol (ralm/retriever/client.py): used by the tests and as a reference."""

    def __init__(self, host: str, port: int, batch_size: int, dim: int, k: int, nprobe: Optional[int] = None):
        self.batch_size, self.dim, self.k, self.nprobe = batch_size, dim, k, nprobe
        self.sock = socket.create_connection((host, port))
        self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)

    def retrieve(self, queries: np.ndarray, list_ids: Optional[np.ndarray] = None):
        msg = encode_request(queries, self.k) if list_ids is None else encode_request_with_lists(queries, list_ids, self.k)
        self.sock.sendall(msg)
        ans = _recv_exact(self.sock, answer_message_len(self.k, self.batch_size))
        if ans is None or len(ans) != answer_message_len(self.k, self.batch_size):
            raise ConnectionError("server closed the connection")
        I, D = decode_answer(ans, self.k, self.batch_size)
        return {"id": I, "dist": D}

    def close(self):
        self.sock.close()
