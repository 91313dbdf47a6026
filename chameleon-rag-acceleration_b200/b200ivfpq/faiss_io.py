"""Faiss index-file container for IndexIVFPQ ("IwPQ", Faiss 1.7.x index_write.cpp / index_read.cpp) -- SURVEY.md
section 8f rank 2: lets the `<db>_<key>_populated.index` files the reference writes with faiss.write_index
(Faiss_experiments/bench_cpu_performance.py:113,164; loaded at llm_inference_gpu/ralm/retriever/
faiss_retriever.py:85-94) enter this engine, and lets an index built here be opened by Faiss.

STATUS: written from the published Faiss 1.7.1 / 1.7.2 serialisation layout; Faiss is not installable in this
environment, so the reader has only been checked against this module's own writer (round trip), NOT against a file
produced by a Faiss binary.  Layout (little endian; size_t and idx_t are 8 bytes, int 4, bool 1):

  "IwPQ"  index header: d:int  ntotal:i64  dummy:i64 (1<<20)  dummy:i64  is_trained:bool  metric_type:int
          nlist:size_t  nprobe:size_t
          quantizer = "IxF2"  index header (d, nlist, ...)  flat storage, one count word + nlist*d*4 payload bytes:
                      count = nlist*d      WRITEVECTOR(xb) up to 1.6.x, and WRITEXBVECTOR(codes) from 1.7.2 on
                                           (codes is the uint8 image of the floats: size/4 = nlist*d) -- the DEFAULT
                                           this module writes;
                      count = nlist*d/4    WRITEXBVECTOR applied to the float vector xb (size/4 ELEMENTS), which is
                                           how we remember 1.7.0 / 1.7.1 -- the release the reference pins for its CPU
                                           runs (Faiss_experiments/README.md:18).  A reviewer of round 1 remembers
                                           WRITEVECTOR there instead; no Faiss source or binary is available here to
                                           settle it, so the reader accepts both counts (the payload is identical and
                                           the count tells them apart) and the writer only emits this one on request
          direct map: type:char  array: n:size_t + n*i64   (type 2 = hashtable: + n:size_t + n*(i64, i64))
          by_residual:bool  code_size:size_t
          ProductQuantizer: d:size_t  M:size_t  nbits:size_t  centroids: n:size_t + n*f32   ((M, ksub, dsub))
          "ilar"  nlist:size_t  code_size:size_t  "full" sizes: n:size_t + n*size_t
                                                | "sprs" pairs (list, size): n:size_t + n*size_t
                  then for every non-empty list: codes (size*code_size bytes), ids (size*i64)
"""
from __future__ import annotations

import struct

import numpy as np

METRIC_L2 = 1


class _R:
    def __init__(self, buf):
        self.b, self.o = memoryview(buf), 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def fourcc(self):
        s = bytes(self.b[self.o:self.o + 4]).decode("ascii", "replace")
        self.o += 4
        return s

    def array(self, dtype, n):
        dt = np.dtype(dtype)
        a = np.frombuffer(self.b, dtype=dt, count=n, offset=self.o)
        self.o += n * dt.itemsize
        return a

    def vector(self, dtype):
        return self.array(dtype, self.take("Q"))


def _read_header(r):
    d, ntotal, _, _, is_trained, metric = r.take("i"), r.take("q"), r.take("q"), r.take("q"), r.take("?"), r.take("i")
    if metric > 1:
        r.take("f")
    return d, ntotal, is_trained, metric


def parse_faiss_ivfpq(path_or_bytes) -> dict:
    """Parse an "IwPQ" file into the flat arrays IndexIVFPQ.set_codebooks / set_lists take.  Pure numpy."""
    buf = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    r = _R(buf)
    h = r.fourcc()
    if h != "IwPQ":
        raise RuntimeError(f"read_index: unsupported index type '{h}' (only IndexIVFPQ 'IwPQ' files are supported)")
    d, ntotal, is_trained, metric = _read_header(r)
    if metric != METRIC_L2:
        raise RuntimeError("read_index: only METRIC_L2 indexes are supported")
    nlist, nprobe = r.take("Q"), r.take("Q")
    hq = r.fourcc()
    if hq != "IxF2":
        raise RuntimeError(f"read_index: unsupported coarse quantizer '{hq}' (only IndexFlatL2)")
    qd, qn, _, _ = _read_header(r)
    n4 = r.take("Q")
    if n4 == qn * qd:                           # count = number of floats (WRITEVECTOR, WRITEXBVECTOR(codes))
        coarse = r.array(np.float32, n4)
    elif n4 * 4 == qn * qd:                     # count = floats / 4 (WRITEXBVECTOR on the float vector)
        coarse = r.array(np.float32, n4 * 4)
    else:
        raise RuntimeError("read_index: unexpected size of the coarse quantizer's vector storage")
    coarse = np.array(coarse, np.float32).reshape(qn, qd)
    dm_type = r.take("b")
    r.vector(np.int64)
    if dm_type == 2:
        r.array(np.int64, 2 * r.take("Q"))
    by_residual, code_size = r.take("?"), r.take("Q")
    pd, M, nbits = r.take("Q"), r.take("Q"), r.take("Q")
    cent = np.array(r.vector(np.float32), np.float32)
    if nbits != 8 or not by_residual or pd != d:
        raise RuntimeError("read_index: only by_residual 8-bit PQ indexes are supported")
    hil = r.fourcc()
    if hil != "ilar":
        raise RuntimeError(f"read_index: unsupported inverted lists '{hil}' (only ArrayInvertedLists)")
    il_nlist, il_code_size = r.take("Q"), r.take("Q")
    kind = r.fourcc()
    sizes = np.zeros(il_nlist, np.int64)
    if kind == "full":
        sizes[:] = r.vector(np.uint64).astype(np.int64)
    elif kind == "sprs":
        pairs = r.vector(np.uint64).astype(np.int64).reshape(-1, 2)
        sizes[pairs[:, 0]] = pairs[:, 1]
    else:
        raise RuntimeError(f"read_index: unknown list encoding '{kind}'")
    offsets = np.zeros(il_nlist + 1, np.int64)
    offsets[1:] = np.cumsum(sizes)
    total = int(offsets[-1])
    codes = np.empty((total, il_code_size), np.uint8)
    ids = np.empty(total, np.int64)
    for l in range(il_nlist):
        n = int(sizes[l])
        if n:
            codes[offsets[l]:offsets[l + 1]] = r.array(np.uint8, n * il_code_size).reshape(n, il_code_size)
            ids[offsets[l]:offsets[l + 1]] = r.array(np.int64, n)
    assert il_nlist == nlist and il_code_size == code_size == M and total == ntotal
    return {"d": d, "nlist": int(nlist), "M": int(M), "nbits": int(nbits), "nprobe": int(nprobe),
            "is_trained": bool(is_trained), "coarse": coarse, "pq": cent.reshape(M, 256, d // M), "offsets": offsets,
            "codes": codes, "ids": ids}


def write_faiss_ivfpq(path, arrays: dict, nprobe: int = 1, flat_storage: str = "bytes") -> None:
    """Write the arrays of IndexIVFPQ.to_arrays() as an "IwPQ" file.  flat_storage: "bytes" (default: count word =
    nlist*d, what Faiss <= 1.6 and >= 1.7.2 read) or "float" (count word = nlist*d / 4, see the module docstring) for the
    coarse quantizer's vectors."""
    if flat_storage not in ("bytes", "float"):
        raise ValueError("flat_storage must be 'bytes' or 'float'")
    coarse = np.ascontiguousarray(arrays["coarse"], np.float32)
    pq = np.ascontiguousarray(arrays["pq"], np.float32)
    offsets = np.ascontiguousarray(arrays["offsets"], np.int64)
    codes = np.ascontiguousarray(arrays["codes"], np.uint8)
    ids = np.ascontiguousarray(arrays["ids"], np.int64)
    nlist, d = coarse.shape
    M = pq.shape[0]
    ntotal = int(offsets[-1])
    out = []

    def header(dd, n):
        out.append(struct.pack("<iqqq?i", dd, n, 1 << 20, 1 << 20, True, METRIC_L2))

    out.append(b"IwPQ")
    header(d, ntotal)
    out.append(struct.pack("<QQ", nlist, nprobe))
    out.append(b"IxF2")
    header(d, nlist)
    if flat_storage == "float":
        assert (nlist * d) % 4 == 0, "Faiss writes the flat storage in groups of 4 elements"
        out.append(struct.pack("<Q", nlist * d // 4))
    else:
        out.append(struct.pack("<Q", nlist * d))
    out.append(coarse.tobytes())
    out.append(struct.pack("<bQ", 0, 0))                                  # no direct map
    out.append(struct.pack("<?Q", True, M))                               # by_residual, code_size
    out.append(struct.pack("<QQQQ", d, M, 8, pq.size))
    out.append(pq.tobytes())
    out.append(b"ilar")
    out.append(struct.pack("<QQ", nlist, M))
    sizes = np.diff(offsets).astype(np.uint64)
    nz = np.nonzero(sizes)[0]
    if len(nz) > nlist // 2:
        out.append(b"full")
        out.append(struct.pack("<Q", nlist))
        out.append(sizes.tobytes())
    else:
        out.append(b"sprs")
        pairs = np.stack([nz.astype(np.uint64), sizes[nz]], 1)
        out.append(struct.pack("<Q", pairs.size))
        out.append(np.ascontiguousarray(pairs).tobytes())
    for l in nz:
        out.append(codes[offsets[l]:offsets[l + 1]].tobytes())
        out.append(ids[offsets[l]:offsets[l + 1]].tobytes())
    with open(path, "wb") as f:
        for chunk in out:
            f.write(chunk)


def read_codebooks_raw(pq_path: str, vq_path: str):
    """The raw dumps of extract_Enzian_U250_required_data.py:510-514:
    product_quantizer_float32_<M>_<256>_<dsub>_raw and vector_quantizer_float32_<nlist>_<d>_raw."""
    def shape_of(p):
        parts = p.rstrip("/").split("/")[-1].split("_")
        return tuple(int(x) for x in parts if x.isdigit())
    pq = np.fromfile(pq_path, np.float32).reshape(shape_of(pq_path)[-3:])
    vq = np.fromfile(vq_path, np.float32).reshape(shape_of(vq_path)[-2:])
    return vq, pq
