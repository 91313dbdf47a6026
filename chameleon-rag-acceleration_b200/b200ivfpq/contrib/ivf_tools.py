"""faiss.contrib.ivf_tools: `from faiss.contrib.ivf_tools import search_preassigned` (ralm/server/faiss_server.py:24,
called at :233 with the list ids the FPGA / client already selected)."""


def search_preassigned(index_ivf, xq, k, list_nos, coarse_dis=None):
    """Search `xq` in the inverted lists `list_nos` (nq, nprobe) int64 -- -1 entries are skipped -- without running the
    coarse quantizer.  `coarse_dis` is accepted for signature compatibility; the residual-LUT distance does not use it."""
    return index_ivf.search_preassigned(xq, k, list_nos)
