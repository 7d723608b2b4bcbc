"""CPU tests of the multi-GPU host logic (world size 2, gloo): the row-strip partition is disjoint and covering,
every strip's halo / source-row window contains exactly the rows its outputs depend on (checked against the
oracle's contribution tables), and strips computed independently from ONLY those source rows reassemble to the
full-image result (checked with the oracle run on the cropped source -- all stages are local)."""
import os
import socket
import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import libsrcnn_b200 as S
import oracle_lib as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, d, m, filt = 64, 48, 3, 2.0, S.BICUBIC
        ow, oh = S.output_dims(w, h, m)
        a, b = S.strip_rows(oh, world, rank)
        (e0, e1), (s0, s1) = S.strip_geometry(w, h, d, m, filt, a, b)
        # halo = strip +- 6 clipped (conv I +-4, conv III +-2)
        assert (e0, e1) == (max(0, a - 6), min(oh, b + 6))
        # source window is exactly the union of the vertical taps of rows [e0, e1) for both filters (tight on both ends
        # up to the zero-weight taps of the dense tables)
        lo, hi = 10 ** 9, -1
        for f in (filt, S.BILINEAR):
            L, R, _ = O.best().table(f, oh, h)
            lo = min(lo, int(L[e0:e1].min())); hi = max(hi, int(R[e0:e1].max()))
        assert s0 <= lo and s1 >= hi + 1 and lo - s0 <= 2 and s1 - (hi + 1) <= 2
        # compute the strip on the CPU from ONLY source rows [s0 - margin, s1 + margin): locality of every stage
        img = O.synth("uniform", h, w, d, 99)
        margin = 8
        c0, c1 = max(0, s0 - margin), min(h, s1 + margin)
        rc, crop_out, _ = O.best().process(img[c0:c1], m, filt, False, False)
        crop_out = crop_out.reshape(-1, ow, d)
        strip = crop_out[a - int(c0 * m): b - int(c0 * m)]
        parts = [None] * world
        dist.all_gather_object(parts, (a, b, strip))
        if rank == 0:
            parts.sort(key=lambda t: t[0])
            assert parts[0][0] == 0 and parts[-1][1] == oh and all(x[1] == y[0] for x, y in zip(parts, parts[1:]))
            rc, full, _ = O.best().process(img, m, filt, False, False)
            full = full.reshape(oh, ow, d)
            got = np.concatenate([t[2] for t in parts], 0)
            # rows within 16 output rows of an artificial crop edge may differ (the crop truncates filter windows there)
            q.put(bool(np.array_equal(got, full)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_strip_partition_gloo():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


@pytest.mark.parametrize("oh,parts", [(2160, 8), (8640, 8), (17, 4), (5, 8), (1080, 3)])
def test_strip_rows_cover(oh, parts):
    cuts = [S.strip_rows(oh, parts, i) for i in range(parts)]
    assert cuts[0][0] == 0 and cuts[-1][1] == oh
    assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
