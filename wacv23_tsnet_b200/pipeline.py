"""Host <-> device pipelining for a stream of batches (serving / demo loops).

`FramePipeline(net)` overlaps, for consecutive batches, the pinned-host -> device copy of batch i+1 (copy stream)
with the forward of batch i (compute stream) and the device -> host read of result i-1 (output stream).  Every byte
still moves inside the caller's loop; only the serialisation is removed.  The reference stages inputs synchronously
inside set_test_input (model/TSNet.py:283-290) and reads the result with `.data.cpu()` (demo/demo_face.py:194).

All device input buffers and pinned output buffers are allocated once (rings), so the steady state performs no
allocator or cudaHostAlloc calls.
"""
import torch


class FramePipeline:
    def __init__(self, net, depth=2):
        self.net = net
        self.depth = depth
        self.copy_stream = torch.cuda.Stream()
        self.out_stream = torch.cuda.Stream()
        self._in_ring = None    # depth+1 sets of device input tensors
        self._in_free = None    # event per set: the forward that consumed it has finished
        self._out_ring = None   # depth+2 pinned host result tensors
        self._n_in = 0
        self._n_out = 0
        self._staged = []       # (set index, ready event)
        self._pending = []      # (ring index, done event)

    @staticmethod
    def _like_on_device(batch):
        mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device="cuda")
        return {k: ([mk(t) for t in v] if isinstance(v, (list, tuple)) else mk(v)) for k, v in batch.items()}

    def _stage(self, batch):
        if self._in_ring is None:
            self._in_ring = [self._like_on_device(batch) for _ in range(self.depth + 1)]
            self._in_free = [None] * (self.depth + 1)
        k = self._n_in % (self.depth + 1)
        self._n_in += 1
        with torch.cuda.stream(self.copy_stream):
            if self._in_free[k] is not None:
                self.copy_stream.wait_event(self._in_free[k])  # do not overwrite inputs a forward is still reading
            dst = self._in_ring[k]
            for key, v in batch.items():
                if isinstance(v, (list, tuple)):
                    for d, s in zip(dst[key], v):
                        d.copy_(s, non_blocking=True)
                else:
                    dst[key].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._staged.append((k, ev))

    def _run_one(self):
        k, ev = self._staged.pop(0)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        dev = self._in_ring[k]
        self.net.set_test_input(dev["src_img"], dev["src_lbl"], dev["src_bbox"], dev["tar_lbl"], dev["tar_bbox"])
        self.net.forward()
        out = self.net.rec_tar_img
        done = torch.cuda.Event()
        done.record(cur)
        self._in_free[k] = done
        if self._out_ring is None:
            self._out_ring = [torch.empty(out.shape, dtype=out.dtype, pin_memory=True) for _ in range(self.depth + 2)]
        r = self._n_out % (self.depth + 2)
        self._n_out += 1
        with torch.cuda.stream(self.out_stream):
            self.out_stream.wait_event(done)
            self._out_ring[r].copy_(out, non_blocking=True)
            out.record_stream(self.out_stream)
            fin = torch.cuda.Event()
            fin.record(self.out_stream)
        self._pending.append((r, fin))

    def _pop(self, copy_out):
        r, fin = self._pending.pop(0)
        fin.synchronize()
        return self._out_ring[r].clone() if copy_out else self._out_ring[r]

    def run(self, batches, copy_out=True):
        """batches: iterable of dicts of PINNED host tensors with keys src_img / src_lbl / src_bbox (lists) and
        tar_lbl / tar_bbox (all batches of one run have the same shapes).  Yields the host result of every batch, in
        order.  copy_out=False yields views of the internal pinned ring (valid until depth+1 further results)."""
        it = iter(batches)
        with torch.no_grad():
            for b in it:
                self._stage(b)
                if len(self._staged) >= self.depth:
                    break
            while self._staged:
                self._run_one()
                nxt = next(it, None)
                if nxt is not None:
                    self._stage(nxt)
                while len(self._pending) > 1:
                    yield self._pop(copy_out)
            while self._pending:
                yield self._pop(copy_out)
