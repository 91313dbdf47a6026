"""Generate tests/golden/wire_format.npz by RUNNING THE REFERENCE'S OWN serialisation code.

Source: /root/reference/Chameleon/llm_inference_gpu/ralm/retriever/serialization_utils.py (imported from where it lies;
this script only runs in the build container).  Seeded inputs go through the reference's encode_request,
encode_request_with_lists and encode_answer; the byte strings it emits, and what its decoders make of them, are stored.
tests/test_server_wire.py::test_wire_bytes_equal_the_references_encoders then requires b200ivfpq.server to emit and
accept exactly those bytes.
"""
import importlib.util
import os

import numpy as np

SRC = "/root/reference/Chameleon/llm_inference_gpu/ralm/retriever/serialization_utils.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_serialization_utils", SRC)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(20240)
    out = {}
    for tag, (batch, dim, nprobe, k) in {"a": (3, 8, 4, 10), "b": (1, 5, 1, 1), "c": (32, 16, 32, 100)}.items():
        q = rng.standard_normal((batch, dim)).astype(np.float32)
        lists = rng.integers(-1, 1 << 20, (batch, nprobe)).astype(np.int64)
        I = rng.integers(-1, 1 << 40, (batch, k)).astype(np.int64)
        D = rng.random((batch, k)).astype(np.float32)
        req = bytes(ref.encode_request(q, k, batch, dim))
        reql = bytes(ref.encode_request_with_lists(q, lists, batch, dim, nprobe, k))
        ans = bytes(ref.encode_answer(I, D, k, batch))
        assert len(req) == ref.request_message_length(batch, dim)
        assert len(reql) == ref.request_message_length_with_lists(batch, dim, nprobe)
        assert len(ans) == ref.answer_message_len(k, batch)
        # what the reference's decoders read back (they must agree with the inputs, or the fixture is useless)
        k2, q2 = ref.decode_request(req, batch, dim)
        k3, q3, l3 = ref.decode_request_with_lists(reql, batch, dim, nprobe)
        I2, D2 = ref.decode_answer(ans, k, batch)
        assert k2 == k and k3 == k
        assert np.array_equal(np.asarray(q2, np.float32).reshape(batch, dim), q)
        assert np.array_equal(np.asarray(q3, np.float32).reshape(batch, dim), q)
        assert np.array_equal(np.asarray(l3, np.int64).reshape(batch, nprobe), lists)
        assert np.array_equal(np.asarray(I2, np.int64).reshape(batch, k), I)
        assert np.array_equal(np.asarray(D2, np.float32).reshape(batch, k), D)
        out.update({f"{tag}_shape": np.array([batch, dim, nprobe, k]), f"{tag}_q": q, f"{tag}_lists": lists,
                    f"{tag}_I": I, f"{tag}_D": D, f"{tag}_req": np.frombuffer(req, np.uint8),
                    f"{tag}_req_lists": np.frombuffer(reql, np.uint8), f"{tag}_ans": np.frombuffer(ans, np.uint8)})
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wire_format.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
