"""The Faiss-style surface end to end on the GPU: train / add / search, accessors, persistence, merge, errors."""
import os

import numpy as np
import pytest

import _util

pytestmark = pytest.mark.gpu


def _clustered(seed, n, d, ncent=64, sigma=0.05):
    rng = np.random.default_rng(seed)
    cent = rng.random((ncent, d), dtype=np.float32)
    return (cent[rng.integers(0, ncent, n)] + rng.standard_normal((n, d)).astype(np.float32) * sigma).astype(np.float32)


def test_train_add_search_like_ivfpq_random_dataset(oracle):
    """The reference's toy driver (IVFPQ_random_dataset.py:20-46) at reduced size, then parity vs the oracle
    on the index it produced, and a recall sanity check against brute force."""
    import b200ivfpq as faiss
    d, nb, nq, nlist, m, k = 64, 40000, 200, 64, 8, 10
    xb, xq = _clustered(1, nb, d), _clustered(2, nq, d)
    quantizer = faiss.IndexFlatL2(d)
    index = faiss.IndexIVFPQ(quantizer, d, nlist, m, 8)
    assert not index.is_trained
    index.train(xb)
    assert index.is_trained and index.quantizer.ntotal == nlist
    for i0 in range(0, nb, 10000):            # add in blocks (bench_cpu_performance.py:155-159)
        index.add(xb[i0:i0 + 10000])
    assert index.ntotal == nb
    index.nprobe = 8
    D, I = index.search(xq, k)
    assert D.shape == (nq, k) and I.shape == (nq, k) and D.dtype == np.float32 and I.dtype == np.int64
    a = index.to_arrays()
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 8, k)
    _util.assert_bit_equal(D, Dr, "D")
    _util.assert_bit_equal(I, Ir, "I")
    # lists are in insertion order with sequential ids
    for l in (0, 5, nlist - 1):
        ids = index.invlists.get_ids(l)
        assert (np.diff(ids) > 0).all()
    # recall@10 vs exact search (IndexFlatL2 over the base): equals the oracle's by id parity; sanity > 0.3
    flat = faiss.IndexFlatL2(d)
    flat.add(xb)
    _, gt = flat.search(xq, k)
    assert oracle.recall_at_k(I, gt, k) == oracle.recall_at_k(Ir, gt, k)
    assert oracle.recall_at_k(I, gt, k) > 0.3


def test_flat_index_exact(oracle):
    import b200ivfpq as faiss
    rng = np.random.default_rng(0)
    xb, xq = rng.random((3000, 48), dtype=np.float32), rng.random((17, 48), dtype=np.float32)
    flat = faiss.index_factory(48, "Flat")
    flat.add(xb)
    D, I = flat.search(xq, 5)
    dr, ir = oracle.C.coarse(xq, xb, 5)
    _util.assert_bit_equal(D, dr)
    _util.assert_bit_equal(I, ir)


def test_accessors_and_persistence(oracle, tmp_path):
    import b200ivfpq as faiss
    a = _util.make_index_arrays(oracle, 3, 32, 10, 8, 1200)
    index = faiss.index_factory(32, "IVF10,PQ8x8")
    index.set_codebooks(a["coarse"], a["pq"])
    index.set_lists(a["offsets"], a["codes"], a["ids"])
    assert (index.d, index.nlist, index.ntotal) == (32, 10, 1200)
    assert (index.pq.M, index.pq.ksub, index.pq.dsub, index.invlists.code_size) == (8, 256, 4, 8)
    cen = index.pq.centroids.reshape(index.pq.M, index.pq.ksub, index.pq.dsub)   # extract script :222-232
    _util.assert_bit_equal(cen, a["pq"])
    _util.assert_bit_equal(index.quantizer.get_xb().reshape(10, 32), a["coarse"])
    for l in range(10):
        ls = index.invlists.list_size(l)
        assert ls == a["offsets"][l + 1] - a["offsets"][l]
        _util.assert_bit_equal(index.invlists.get_ids(l), a["ids"][a["offsets"][l]:a["offsets"][l + 1]])
        _util.assert_bit_equal(index.invlists.get_codes(l).reshape(ls, 8),
                               a["codes"][a["offsets"][l]:a["offsets"][l + 1]])
    ps = faiss.ParameterSpace()
    ps.initialize(index)
    ps.set_index_parameters(index, "nprobe=4")
    assert index.nprobe == 4
    xq = _util.make_queries(1, a, 9)
    D, I = index.search(xq, 5)
    fn = os.path.join(tmp_path, "toy_IVF10,PQ8_populated.index")
    faiss.write_index(index, fn)
    index2 = faiss.read_index(fn)
    assert index2.nprobe == 4 and index2.ntotal == 1200
    D2, I2 = index2.search(xq, 5)
    _util.assert_bit_equal(D2, D)
    _util.assert_bit_equal(I2, I)


def test_retriever_adapter(oracle):
    """LocalFaissRetriever.retrieve contract (faiss_retriever.py:227-275) and IndexScanner (index_scanner.py)."""
    import b200ivfpq as faiss
    a = _util.make_index_arrays(oracle, 13, 64, 16, 16, 4000)
    index = faiss.IndexIVFPQ(faiss.IndexFlatL2(64), 64, 16, 16, 8)
    index.set_codebooks(a["coarse"], a["pq"])
    index.set_lists(a["offsets"], a["codes"], a["ids"])
    r = faiss.LocalB200Retriever(index, default_k=10, nprobe=4)
    xq = _util.make_queries(3, a, 6)
    out = r.retrieve(xq, nprobe=4, k=10)
    assert set(out) == {"id", "dist"} and out["id"].dtype == np.int64 and out["dist"].dtype == np.float32
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 4, 10)
    _util.assert_bit_equal(out["id"], Ir)
    _util.assert_bit_equal(out["dist"], Dr)
    scanner = faiss.IndexScanner(dim=64, nlist=16, nprobe=4, centroids=a["coarse"])
    list_ids, list_cent = scanner.search(xq)
    _, pid = oracle.C.coarse(xq, a["coarse"], 4)
    _util.assert_bit_equal(list_ids, pid)
    assert list_cent.shape == (6, 4, 64)
    out2 = r.retrieve_with_lists(xq, list_ids, k=10)
    _util.assert_bit_equal(out2["id"], Ir)


def test_errors_are_runtime_errors(oracle):
    import b200ivfpq as faiss
    index = faiss.index_factory(32, "IVF8,PQ4")
    xq = np.zeros((2, 32), np.float32)
    with pytest.raises(RuntimeError):
        index.search(xq, 5)                       # not trained
    a = _util.make_index_arrays(oracle, 1, 32, 8, 4, 100)
    index.set_codebooks(a["coarse"], a["pq"])
    index.set_lists(a["offsets"], a["codes"], a["ids"])
    with pytest.raises(RuntimeError):
        index.search(xq, 0)
    with pytest.raises(RuntimeError):
        index.search(xq, 5000)
    with pytest.raises(AssertionError):
        index.search(np.zeros((2, 31), np.float32), 5)
    with pytest.raises(TypeError):
        index.search(np.zeros((2, 32), np.float64), 5)
    index.nprobe = 100                             # > nlist: clamped like Faiss
    D, I = index.search(xq, 5)
    assert D.shape == (2, 5)
    lib = faiss.load_library()
    assert lib.b200_ivfpq_search(None, 1, None, 1, 1, None, None, None) != 0
    assert b"null" in lib.b200_ivfpq_last_error()


def test_merge_shards_kernel_and_sharded_search(oracle):
    """K5 vs the oracle's merge, and shard-by-position search == unsharded search (modulo ties)."""
    import torch
    import b200ivfpq as faiss
    a = _util.make_index_arrays(oracle, 17, 128, 32, 16, 12000, id_scramble=False)
    xq = _util.make_queries(6, a, 40)
    index = faiss.IndexIVFPQ(faiss.IndexFlatL2(128), 128, 32, 16, 8)
    index.set_codebooks(a["coarse"], a["pq"])
    index.set_lists(a["offsets"], a["codes"], a["ids"])
    index.nprobe = 8
    D, I = index.search(xq, 10)
    for world in (2, 4):
        Ds, Is = [], []
        for r in range(world):
            sub = faiss.shard_index(index, r, world)
            assert abs(sub.ntotal - 12000 / world) <= 1
            d_, i_ = sub.search(torch.from_numpy(xq).cuda(), 10)
            Ds.append(d_)
            Is.append(i_)
        Ds, Is = torch.stack(Ds), torch.stack(Is)
        Dm, Im = faiss.merge_shards(Ds, Is)
        Dm, Im = Dm.cpu().numpy(), Im.cpu().numpy()
        Do, Io = oracle.C.merge_shards(Ds.cpu().numpy(), Is.cpu().numpy())
        _util.assert_bit_equal(Dm, Do, "K5 vs oracle merge D")
        _util.assert_bit_equal(Im, Io, "K5 vs oracle merge I")
        _util.assert_same_modulo_ties(Dm, Im, D, I, f"sharded x{world} vs single")
    # whole lists per shard (shard_mode="list"): the probes of the other shards' lists are masked out; since every
    # (query, list) pair is evaluated exactly once, by the same code, the merged DISTANCES are bit-identical
    xq_t = torch.from_numpy(xq).cuda()
    _, probes = index.quantizer.search(xq_t, 8)
    for world in (2, 3):
        Ds, Is, tot = [], [], 0
        for r in range(world):
            sub = faiss.shard_index_by_list(index, r, world)
            tot += sub.ntotal
            sizes = np.diff(sub._offsets)
            assert not sizes[np.arange(32) % world != r].any()
            masked = torch.where(probes % world == r, probes, torch.full_like(probes, -1))
            d_, i_ = sub.search_preassigned(xq_t, 10, masked)
            Ds.append(d_)
            Is.append(i_)
        assert tot == 12000
        Dm, Im = faiss.merge_shards(torch.stack(Ds), torch.stack(Is))
        _util.assert_bit_equal(Dm.cpu().numpy(), np.asarray(D), f"list-sharded x{world} D")
        _util.assert_same_modulo_ties(Dm.cpu().numpy(), Im.cpu().numpy(), D, I, f"list-sharded x{world} vs single")
    # pack / unpack used by the all-gather
    from b200ivfpq.shards import pack_results, unpack_results
    d2, i2 = unpack_results(pack_results(Ds[0], Is[0]))
    assert torch.equal(d2, Ds[0]) and torch.equal(i2, Is[0])


def test_async_retriever_overlaps_a_decode_stream(oracle):
    """retrieve_send / poll / retrieve_recv (retriever.py:109-163 protocol) on a side stream: the query is a CUDA
    tensor produced on the caller's stream right before the send (like the decoder's hidden state), other work is
    enqueued while the retrieval runs, and the answer matches the oracle."""
    import torch
    import b200ivfpq as faiss
    a = _util.make_index_arrays(oracle, 41, 128, 32, 16, 20000)
    index = faiss.IndexIVFPQ(faiss.IndexFlatL2(128), 128, 32, 16, 8)
    index.set_codebooks(a["coarse"], a["pq"])
    index.set_lists(a["offsets"], a["codes"], a["ids"])
    r = faiss.AsyncB200Retriever(index, default_k=10, nprobe=6)
    xq = _util.make_queries(9, a, 4)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 6, 10)
    for step in range(3):
        hidden = torch.from_numpy(xq).cuda() * 1.0          # produced on the current stream just before the send
        r.retrieve_send(hidden, k=10)
        w = torch.randn(2048, 2048, device="cuda")
        y = w @ w                                            # "decode" work enqueued while the retrieval runs
        with pytest.raises(RuntimeError):
            r.retrieve_send(hidden, k=10)                    # one outstanding request, like the socket protocol
        I, D = r.retrieve_recv(10)
        torch.cuda.synchronize()
        assert r.poll()
        _util.assert_bit_equal(D.cpu().numpy(), Dr, "D")
        _util.assert_bit_equal(I.cpu().numpy(), Ir, "I")
        assert y.shape == (2048, 2048)
    out = r.retrieve(xq, nprobe=6, k=10)                     # numpy in -> numpy out
    _util.assert_bit_equal(out["id"], Ir)
    _util.assert_bit_equal(out["dist"], Dr)
    with pytest.raises(RuntimeError):
        r.retrieve_recv(10)
