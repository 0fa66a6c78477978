"""Request coalescing and GPU dispatch for the A1111-compatible server (reference caller:
flux_app.py:90-204, which generates one request at a time on one device).

The hot path is embarrassingly parallel over images, and one B200 is most efficient at 8 images of one
shape per forward.  So concurrent requests are coalesced:

  * requests with the same (model, height, width, steps, guidance) are compatible -- their images can share a batch
    even when the prompts differ: the pipeline conditions every image on its own prompt (flux/flux.py
    `_denoising_loop`: per-row CLIP vector / modulation when the rows differ) and every image's prior is keyed by
    (its request's seed, its index inside the request), so an image is bit-identical to what the request would have
    produced alone (tests/test_gpu_flux_app.py);
  * the images of a compatible group are cut into jobs of at most `max_batch` (8) images, in arrival order, and the
    jobs are handed to the least-loaded worker -- one worker per GPU (a process pinned to cuda:i, or, for a single
    device, a thread in the server process).  No collective is involved: a worker owns a full replica of the weights.

`coalesce()` is a pure function (unit-tested on the CPU); `NodeScheduler` owns the dispatcher thread and the workers.
"""
from __future__ import annotations

import itertools
import queue
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

MAX_BATCH = 8          # images per forward on one GPU (BASELINE configs[3]: 64 images / 8 GPUs)
COALESCE_WINDOW_S = 0.01


@dataclass
class ImageRequest:
    """One txt2img request as the scheduler sees it (flux_app.py:123-160 arguments)."""
    prompt: str
    model: str = "schnell"
    height: int = 512
    width: int = 512
    steps: Optional[int] = None
    guidance: float = 4.0
    seed: Optional[int] = None
    n_images: int = 1
    rid: int = field(default_factory=itertools.count().__next__)

    def key(self) -> tuple:
        return (self.model, self.height, self.width, self.steps, float(self.guidance))


@dataclass
class Job:
    """At most MAX_BATCH images of one shape: `items` = (request id, image index inside its request, prompt, seed)."""
    key: tuple
    items: List[Tuple[int, int, str, Optional[int]]]


def coalesce(requests: Sequence[ImageRequest], max_batch: int = MAX_BATCH) -> List[Job]:
    """Group compatible requests (same key) in arrival order and cut each group into jobs of <= max_batch images.
    Images of one request stay in order; a request larger than max_batch spans several jobs."""
    groups: Dict[tuple, List[Tuple[int, int, str, Optional[int]]]] = {}
    order: List[tuple] = []
    for r in requests:
        if r.n_images <= 0:
            continue
        k = r.key()
        if k not in groups:
            groups[k] = []
            order.append(k)
        groups[k] += [(r.rid, i, r.prompt, r.seed) for i in range(r.n_images)]
    jobs = []
    for k in order:
        items = groups[k]
        jobs += [Job(k, items[i:i + max_batch]) for i in range(0, len(items), max_batch)]
    return jobs


class _Pending:
    def __init__(self, req: ImageRequest):
        self.req = req
        self.images: List[Optional[object]] = [None] * req.n_images
        self.left = req.n_images
        self.error: Optional[BaseException] = None
        self.done = threading.Event()


class NodeScheduler:
    """Dispatcher thread + workers.  `workers` are callables `run(job) -> list of images` (one per job item, in order),
    one per GPU; each is driven by its own thread, so a worker that is a proxy for a per-GPU process simply blocks on
    that process's pipe.  `submit()` blocks until every image of the request is there and returns them in order."""

    def __init__(self, workers: Sequence[Callable[[Job], list]], max_batch: int = MAX_BATCH,
                 window_s: float = COALESCE_WINDOW_S):
        if not workers:
            raise ValueError("NodeScheduler needs at least one worker")
        self.max_batch, self.window_s = max_batch, window_s
        self._inbox: "queue.Queue[_Pending]" = queue.Queue()
        self._pending: Dict[int, _Pending] = {}
        self._lock = threading.Lock()
        self._queues = [queue.Queue() for _ in workers]
        self._load = [0] * len(workers)   # images queued or running per worker
        self.stats = {"jobs": 0, "images": 0, "batches": []}
        self._stop = False
        self._threads = [threading.Thread(target=self._dispatch, daemon=True)]
        self._threads += [threading.Thread(target=self._work, args=(i, w), daemon=True) for i, w in enumerate(workers)]
        for t in self._threads:
            t.start()

    # ---------------------------------------------------------------- client side
    def submit(self, req: ImageRequest, timeout: Optional[float] = None) -> list:
        p = _Pending(req)
        if req.n_images <= 0:
            return []
        with self._lock:
            self._pending[req.rid] = p
        self._inbox.put(p)
        if not p.done.wait(timeout):
            raise TimeoutError(f"request {req.rid} timed out")
        if p.error is not None:
            raise p.error
        return p.images

    def close(self):
        self._stop = True
        self._inbox.put(None)
        for q in self._queues:
            q.put(None)

    # ---------------------------------------------------------------- dispatcher
    def _dispatch(self):
        while not self._stop:
            first = self._inbox.get()
            if first is None:
                return
            batch = [first]
            deadline = time.monotonic() + self.window_s   # short window: requests arriving together share a forward
            while True:
                left = deadline - time.monotonic()
                try:
                    nxt = self._inbox.get(timeout=max(left, 0.0)) if left > 0 else self._inbox.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    return
                batch.append(nxt)
            for job in coalesce([p.req for p in batch], self.max_batch):
                with self._lock:
                    w = min(range(len(self._load)), key=lambda i: self._load[i])   # least-loaded GPU
                    self._load[w] += len(job.items)
                    self.stats["jobs"] += 1
                    self.stats["images"] += len(job.items)
                    self.stats["batches"].append((w, len(job.items)))
                self._queues[w].put(job)

    # ---------------------------------------------------------------- workers
    def _work(self, idx: int, run: Callable[[Job], list]):
        q = self._queues[idx]
        while True:
            job = q.get()
            if job is None:
                return
            try:
                images = run(job)
                if len(images) != len(job.items):
                    raise RuntimeError(f"worker {idx} returned {len(images)} images for {len(job.items)} items")
                err = None
            except BaseException as e:  # noqa: BLE001  (reported to every request of the job: HTTP 500 upstream)
                images, err = [None] * len(job.items), e
            with self._lock:
                self._load[idx] -= len(job.items)
                for (rid, i, _, _), img in zip(job.items, images):
                    p = self._pending.get(rid)
                    if p is None:
                        continue
                    if err is not None:
                        p.error = err
                    p.images[i] = img
                    p.left -= 1
                    if p.left == 0 or err is not None:
                        self._pending.pop(rid, None)
                        p.done.set()


# ---------------------------------------------------------------------------------------------
# per-GPU worker processes (a full weight replica each; no collective on the data path)
# ---------------------------------------------------------------------------------------------
def _gpu_worker_main(device: str, make_runner, conn):
    """Entry point of a worker process: builds its runner on `device` and serves jobs from the pipe."""
    try:
        run = make_runner(device)
        conn.send(("ready", None))
    except BaseException as e:  # noqa: BLE001
        conn.send(("error", repr(e)))
        return
    while True:
        job = conn.recv()
        if job is None:
            return
        try:
            conn.send(("ok", run(job)))
        except BaseException as e:  # noqa: BLE001
            conn.send(("error", repr(e)))


class ProcessWorker:
    """Proxy for one worker process pinned to one GPU: callable like an in-process worker."""

    def __init__(self, device: str, make_runner):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")   # CUDA contexts do not survive fork
        self.conn, child = ctx.Pipe()
        self.proc = ctx.Process(target=_gpu_worker_main, args=(device, make_runner, child), daemon=True)
        self.proc.start()
        self._ready = False

    def __call__(self, job: Job) -> list:
        if not self._ready:
            kind, val = self.conn.recv()
            if kind != "ready":
                raise RuntimeError(f"worker failed to start: {val}")
            self._ready = True
        self.conn.send(job)
        kind, val = self.conn.recv()
        if kind != "ok":
            raise RuntimeError(val)
        return val

    def close(self):
        try:
            self.conn.send(None)
        except Exception:  # noqa: BLE001
            pass
