"""Request coalescing and dispatch (flux_serve.py) on the CPU with fake workers: grouping rules,
batch limits, routing of images back to their requests, least-loaded placement, error propagation."""
import threading
import time

import pytest

from flux_serve import ImageRequest, Job, NodeScheduler, coalesce


def R(prompt, n=1, **kw):
    return ImageRequest(prompt=prompt, n_images=n, **kw)


def test_coalesce_groups_compatible_requests_in_arrival_order():
    a, b, c = R("a", 3, seed=1), R("b", 2, seed=2, height=1024, width=1024), R("c", 4, seed=3)
    d = R("d", 2, seed=4, steps=4)                       # different steps: its own group
    jobs = coalesce([a, b, c, d], max_batch=8)
    assert [len(j.items) for j in jobs] == [7, 2, 2]
    assert [it[:2] for it in jobs[0].items] == [(a.rid, 0), (a.rid, 1), (a.rid, 2), (c.rid, 0), (c.rid, 1), (c.rid, 2), (c.rid, 3)]
    assert jobs[0].key == a.key() == c.key() and jobs[1].key == b.key() and jobs[2].key == d.key()
    assert [it[2] for it in jobs[0].items] == ["a"] * 3 + ["c"] * 4 and [it[3] for it in jobs[0].items] == [1] * 3 + [3] * 4


def test_coalesce_cuts_groups_at_the_per_gpu_batch():
    big = R("x", 19, seed=7)
    jobs = coalesce([big, R("y", 2, seed=8)], max_batch=8)
    assert [len(j.items) for j in jobs] == [8, 8, 5]
    idx = [it[1] for j in jobs for it in j.items if it[0] == big.rid]
    assert idx == list(range(19))                         # a request's images stay in order across jobs
    assert coalesce([R("z", 0)]) == [] and coalesce([]) == []
    assert R("p", guidance=4).key() == R("q", guidance=4.0).key() != R("q", guidance=3.5).key()


def test_scheduler_routes_images_back_and_balances_gpus():
    seen = [[], []]

    def worker(i):
        def run(job: Job):
            seen[i].append(len(job.items))
            time.sleep(0.02)
            return [f"{rid}:{k}:{prompt}:{seed}" for rid, k, prompt, seed in job.items]
        return run

    s = NodeScheduler([worker(0), worker(1)], max_batch=4, window_s=0.05)
    reqs = [R(f"p{i}", 3, seed=i) for i in range(6)]     # 18 images of one shape -> 4+4+4+4+2 over two workers
    out = [None] * len(reqs)

    def client(i):
        out[i] = s.submit(reqs[i], timeout=10)

    ts = [threading.Thread(target=client, args=(i,)) for i in range(len(reqs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for r, imgs in zip(reqs, out):
        assert imgs == [f"{r.rid}:{k}:{r.prompt}:{r.seed}" for k in range(3)]
    assert s.stats["images"] == 18 and sum(map(sum, seen)) == 18
    assert max(n for w in seen for n in w) <= 4 and seen[0] and seen[1]        # both GPUs worked, batches <= max_batch
    assert s.stats["jobs"] < 18                                                  # requests did share forwards
    s.close()


def test_scheduler_propagates_worker_errors():
    def bad(job):
        raise RuntimeError("boom")

    s = NodeScheduler([bad])
    with pytest.raises(RuntimeError, match="boom"):
        s.submit(R("x", 2, seed=1), timeout=5)
    assert s.submit(R("none", 0)) == []
    with pytest.raises(ValueError):
        NodeScheduler([])
    s.close()
