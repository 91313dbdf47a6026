"""The slice of faiss.contrib the reference imports on the search path (faiss_server.py:24)."""
