"""Index persistence: the checkpoint/resume role of faiss.write_index / read_index
(bench_cpu_performance.py:113,164) for this engine, as one .npz of the flat arrays the reference's
extraction scripts dump (extract_Enzian_U250_required_data.py:222-279, 510-564).
Reading Faiss's own binary format is listed as "next" (SURVEY.md section 8f, rank 2)."""
from __future__ import annotations

import numpy as np

from .faiss_io import parse_faiss_ivfpq, write_faiss_ivfpq
from .index import IndexFlatL2, IndexIVFPQ


def _is_faiss_file(fname: str) -> bool:
    try:
        with open(fname, "rb") as f:
            return f.read(4) == b"IwPQ"
    except OSError:
        return False


def write_index(index: IndexIVFPQ, fname: str) -> None:
    """`.index` / `.faiss` names are written in Faiss's IwPQ container (faiss_io.py), anything else as .npz."""
    if fname.endswith((".index", ".faiss")):
        write_faiss_ivfpq(fname, index.to_arrays(), nprobe=index.nprobe)
        return
    a = index.to_arrays() if index.is_trained else {}
    np.savez(fname if fname.endswith(".npz") else fname + ".npz", d=index.d, nlist=index.nlist, M=index.pq.M,
             nbits=index.pq.nbits, nprobe=index.nprobe, is_trained=index.is_trained, **a)


# faiss.read_index(fname, io_flags) (bench_cpu_performance.py:58,116,167): the flags choose how Faiss maps the file; the
# lists live in HBM here, so they are accepted and have nothing to select
IO_FLAG_MMAP = 1
IO_FLAG_READ_ONLY = 2
IO_FLAG_ONDISK_SAME_DIR = 4


def read_index(fname: str, io_flags: int = 0) -> IndexIVFPQ:
    if _is_faiss_file(fname):
        z = parse_faiss_ivfpq(fname)
        index = IndexIVFPQ(IndexFlatL2(z["d"]), z["d"], z["nlist"], z["M"], z["nbits"])
        index.nprobe = z["nprobe"]
        index.set_codebooks(z["coarse"], z["pq"])
        index.set_lists(z["offsets"], z["codes"], z["ids"])
        return index
    z = np.load(fname if fname.endswith(".npz") else fname + ".npz")
    index = IndexIVFPQ(IndexFlatL2(int(z["d"])), int(z["d"]), int(z["nlist"]), int(z["M"]), int(z["nbits"]))
    index.nprobe = int(z["nprobe"])
    if bool(z["is_trained"]):
        index.set_codebooks(z["coarse"], z["pq"])
        index.set_lists(z["offsets"], z["codes"], z["ids"])
    return index
