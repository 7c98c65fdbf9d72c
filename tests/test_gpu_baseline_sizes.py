"""GPU parity at the sizes the bench numbers are quoted on (BASELINE.json configs[1] and configs[3]), through the C ABI.

* config 4 (1M database x 100k queries): the full-size scan, a 2 000-query sample of it compared field by field
  with the oracle's scan of the whole database (src/ORBmatcher.cpp:37-67), and the capacity sweep over query
  counts below the handle's maximum (the segment count grows when the query count shrinks).
* config 2 (256-frame VGA batch): the exact arrangements `bench.py` times -- one device call split in two halves
  on two streams, and two handles alternating on two streams -- with 32 frames spread over the batch compared
  field by field with the oracle (src/ORBextractor.cpp:1001-1065).
"""
import numpy as np
import pytest

from orbslam_in_practice_b200.synth import synth_batch, synth_descriptor_db, synth_queries

pytestmark = pytest.mark.gpu

ANGLE_TOL_DEG = 1e-3 * 180.0 / np.pi   # 1e-3 rad (BASELINE.json north_star)


def _ncores():
    import os
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def test_config4_full_db_2000_query_sample_vs_oracle(orbx, oracle):
    import torch
    ndb, nq = 1_000_000, 100_000
    db = synth_descriptor_db(ndb); q = synth_queries(db, nq)
    m = orbx.Matcher(nq, ndb)
    dev = torch.device("cuda:0")
    tq, tdb = torch.from_numpy(q).to(dev), torch.from_numpy(db).to(dev)
    out = torch.empty((4, nq), dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    m.knn2_device(tq.data_ptr(), nq, tdb.data_ptr(), ndb, 0, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), s)
    m.ratio_select_device(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), nq, 50, 0.7, out[3].data_ptr(), s)
    torch.cuda.synchronize()
    d1, i1, d2, mt = (out[i].cpu().numpy() for i in range(4))
    # every 50th query: 2 000 queries x 1M rows = 2e9 pairs for the oracle (about a second on the box's cores)
    sel = np.arange(0, nq, 50)
    od1, oi1, od2 = oracle.knn2(q[sel], db, 0, _ncores())
    assert np.array_equal(d1[sel], od1), "best distance"
    assert np.array_equal(i1[sel], oi1), "best index (first minimal row wins)"
    assert np.array_equal(d2[sel], od2), "second-best distance"
    assert np.array_equal(mt[sel], oracle.ratio_select(od1, oi1, od2, 50, 0.7)), "TH_LOW / ratio acceptance"
    # the database holds 1 % duplicated rows: some sampled queries must actually hit a tie (d1 == d2)
    assert (od1 == od2).any()


@pytest.mark.parametrize("max_q,max_db", [(1500, 300_000), (4500, 1_000_000)])
def test_knn_capacity_holds_for_every_query_count(orbx, oracle, max_q, max_db):
    """A handle created for (max_q, max_db) must accept every nq <= max_q at ndb = max_db (the number of
    database segments grows when the query count shrinks, and the partial buffer was sized for nq = max_q only)."""
    import torch
    db = synth_descriptor_db(max_db, seed=99); q = synth_queries(db, max_q, seed=98)
    m = orbx.Matcher(max_q, max_db)
    dev = torch.device("cuda:0")
    tq, tdb = torch.from_numpy(q).to(dev), torch.from_numpy(db).to(dev)
    out = torch.empty((3, max_q), dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    sweep = sorted(set([1, 31, 255, 256, 257, 511, 1000, 1023, 1024, 1025, 1900, 2047, 2048, 2400, 3000, max_q - 1, max_q]) & set(range(1, max_q + 1)))
    for nq in sweep:
        m.knn2_device(tq.data_ptr(), nq, tdb.data_ptr(), max_db, 0, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), s)
    torch.cuda.synchronize()
    # the last sweep entry ran nq = max_q: check a sample of it against the oracle
    sel = np.arange(0, max_q, max(1, max_q // 64))
    od1, oi1, od2 = oracle.knn2(q[sel], db, 0, _ncores())
    got = out.cpu().numpy()
    assert np.array_equal(got[0][sel], od1) and np.array_equal(got[1][sel], oi1) and np.array_equal(got[2][sel], od2)


def _check_frames(oracle, imgs, frames, kps, desc, counts):
    oex = oracle.OracleExtractor()
    total = 0
    for f in frames:
        ko, do = oex(imgs[f])
        n = int(counts[f])
        assert n == len(ko), "frame %d: keypoint count gpu %d oracle %d" % (f, n, len(ko))
        kg, dg = kps[f, :n], desc[f, :n]
        for fld in ("x", "y", "size", "response", "octave", "class_id"):
            assert np.array_equal(kg[fld], ko[fld]), "frame %d: keypoint field %s differs" % (f, fld)
        d = np.abs(kg["angle"] - ko["angle"]); d = np.minimum(d, 360.0 - d)
        assert d.max() <= ANGLE_TOL_DEG, "frame %d: angle error %g deg" % (f, d.max())
        bits = int(np.unpackbits(dg ^ do).sum())
        assert bits <= 1e-3 * dg.size * 8, "frame %d: descriptor bits differing: %d" % (f, bits)
        total += n
    return total


def test_config2_256_frame_batch_vs_oracle(orbx, oracle):
    """The two arrangements bench.py's `value` and `single_handle` time, at the full batch of 256 frames."""
    import torch
    B, W, H = 256, 640, 480
    nuniq = 64
    base = synth_batch(range(1000, 1000 + nuniq), W, H)
    # every frame of the batch is distinct from its neighbours: frame f = base[(7 f) mod 64] (7 is coprime to 64)
    order = (7 * np.arange(B)) % nuniq
    imgs = np.ascontiguousarray(base[order])
    dev = torch.device("cuda:0")
    d_frames = torch.from_numpy(imgs).to(dev)
    exA = orbx.Extractor(max_width=W, max_height=H, max_batch=B)
    exB = orbx.Extractor(max_width=W, max_height=H, max_batch=B)
    cap = exA.capacity
    sA, sB = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def outputs():
        return (torch.zeros((B, cap, 7), dtype=torch.float32, device=dev), torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev),
                torch.full((B,), -1, dtype=torch.int32, device=dev))

    def host(o):
        k = o[0].cpu().numpy().view(np.float32).reshape(B, cap, 7)
        kp = np.zeros((B, cap), orbx.KEYPOINT_DTYPE)
        kp.view(np.uint8).reshape(B, cap, 28)[:] = k.view(np.uint8).reshape(B, cap, 28)
        return kp, o[1].cpu().numpy(), o[2].cpu().numpy()

    frames = sorted(set(list(range(0, B, 8)) + [127, 128, 129, 255]))       # 32 frames over the batch + the split seam
    assert len(frames) >= 32

    # (1) one handle, the batch split in two independent halves on two streams (the library default for >= 32 frames)
    exA.set_device_split(2)
    o1 = outputs()
    exA.extract_device(d_frames.data_ptr(), W, W * H, W, H, B, o1[0].data_ptr(), o1[1].data_ptr(), o1[2].data_ptr(), sA.cuda_stream)
    sA.synchronize()
    kp1, de1, cn1 = host(o1)
    total = _check_frames(oracle, imgs, frames, kp1, de1, cn1)
    assert total > 900 * len(frames)

    # (2) two handles alternating on two streams, several steps in flight (bench.py's step_pipelined), unsplit
    exA.set_device_split(1); exB.set_device_split(1)
    oA, oB = outputs(), outputs()
    for i in range(4):
        e, o, s = (exB, oB, sB) if i & 1 else (exA, oA, sA)
        e.extract_device(d_frames.data_ptr(), W, W * H, W, H, B, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    for o in (oA, oB):
        kp, de, cn = host(o)
        assert np.array_equal(cn, cn1), "per-frame counts differ between the arrangements"
        # all 256 frames identical to arrangement (1), which was compared with the oracle on the sample
        for f in range(B):
            n = int(cn[f])
            assert kp[f, :n].tobytes() == kp1[f, :n].tobytes() and np.array_equal(de[f, :n], de1[f, :n]), "frame %d" % f
    # frames that repeat a base image inside the batch must give identical results wherever they sit
    for f in range(nuniq, B):
        g = f - nuniq
        n = int(cn1[f])
        assert cn1[g] == n and kp1[f, :n].tobytes() == kp1[g, :n].tobytes() and np.array_equal(de1[f, :n], de1[g, :n])


def test_config2_host_pipeline_256_frames_vs_oracle(orbx, oracle):
    """The e2e arrangement: orbx_extract_host_begin/_end on pinned host buffers, 256 frames, chunks over the copy streams."""
    import torch
    B, W, H = 256, 640, 480
    nuniq = 32
    base = synth_batch(range(2000, 2000 + nuniq), W, H)
    order = (5 * np.arange(B)) % nuniq
    imgs = np.ascontiguousarray(base[order])
    ex = orbx.Extractor(max_width=W, max_height=H, max_batch=B)
    cap = ex.capacity
    hf = torch.from_numpy(imgs).pin_memory()
    hk = torch.zeros((B, cap, 7), dtype=torch.float32).pin_memory()
    hd = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
    hc = torch.full((B,), -1, dtype=torch.int32).pin_memory()
    ex.extract_host_begin(hf.data_ptr(), W, W * H, W, H, B, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
    ex.extract_host_end()
    kp = np.zeros((B, cap), orbx.KEYPOINT_DTYPE)
    kp.view(np.uint8).reshape(B, cap, 28)[:] = hk.numpy().view(np.uint8).reshape(B, cap, 28)
    frames = list(range(0, B, 8)) + [31, 33, 63, 65, 95, 97, 191, 193, 223, 225, 255]     # incl. both sides of the seams of the eight host chunks
    _check_frames(oracle, imgs, frames, kp, hd.numpy(), hc.numpy())


@pytest.mark.parametrize("nframes,chunks", [(37, 8), (37, 5), (70, 3), (9, 8)])
def test_host_chunk_seams_with_ragged_batches(orbx, oracle, nframes, chunks, monkeypatch):
    """orbx_extract_host with frame counts that do not divide into the chunk count (ORBX_HOST_CHUNKS is read per call): every
    frame next to a chunk seam against the oracle, counts of all frames against a one-chunk run."""
    W, H = 320, 240
    imgs = np.ascontiguousarray(synth_batch(range(3000, 3000 + nframes), W, H))
    ex = orbx.Extractor(max_width=W, max_height=H, max_batch=nframes)
    monkeypatch.setenv("ORBX_HOST_CHUNKS", "1")
    k1, d1, c1 = ex.extract_host(imgs)
    monkeypatch.setenv("ORBX_HOST_CHUNKS", str(chunks))
    kp, desc, counts = ex.extract_host(imgs)
    assert np.array_equal(c1, counts)
    for f in range(nframes):
        n = int(counts[f])
        assert np.array_equal(k1[f, :n], kp[f, :n]) and np.array_equal(d1[f, :n], desc[f, :n]), "frame %d differs between 1 and %d chunks" % (f, chunks)
    seams = sorted({min(nframes - 1, max(0, nframes * c // chunks + d)) for c in range(chunks + 1) for d in (-1, 0)})
    _check_frames(oracle, imgs, seams, kp, desc, counts)
