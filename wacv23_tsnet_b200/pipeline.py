"""Host <-> device pipelining for a stream of batches (serving / demo loops).

`FramePipeline(net)` overlaps, for consecutive batches, the pinned-host -> device copy of batch i+1 (copy stream)
with the forward of batch i (compute stream) and the device -> host read of result i-1.  Every byte still moves
inside the caller's loop; only the serialisation is removed.  The reference stages inputs synchronously inside
set_test_input (model/TSNet.py:283-290) and reads the result with `.data.cpu()` (demo/demo_face.py:194).
"""
import torch


class FramePipeline:
    def __init__(self, net, depth=2):
        self.net = net
        self.copy_stream = torch.cuda.Stream()
        self.out_stream = torch.cuda.Stream()
        self.depth = depth
        self._staged = []     # (device input dict, ready event)
        self._pending = []    # (pinned host result, done event)

    @staticmethod
    def _to_device(batch):
        mv = lambda t: t.cuda(non_blocking=True)
        return {k: ([mv(t) for t in v] if isinstance(v, (list, tuple)) else mv(v)) for k, v in batch.items()}

    def _stage(self, batch):
        with torch.cuda.stream(self.copy_stream):
            dev = self._to_device(batch)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._staged.append((dev, ev))

    def _run_one(self):
        dev, ev = self._staged.pop(0)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        self.net.set_test_input(dev["src_img"], dev["src_lbl"], dev["src_bbox"], dev["tar_lbl"], dev["tar_bbox"])
        self.net.forward()
        out = self.net.rec_tar_img
        done = torch.cuda.Event()
        done.record(cur)
        for v in dev.values():  # the copy stream allocated these tensors; tell the allocator who else used them
            for t in (v if isinstance(v, (list, tuple)) else [v]):
                t.record_stream(cur)
        with torch.cuda.stream(self.out_stream):
            self.out_stream.wait_event(done)
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            out.record_stream(self.out_stream)
            fin = torch.cuda.Event()
            fin.record(self.out_stream)
        self._pending.append((host, fin))

    def run(self, batches):
        """batches: iterable of dicts of PINNED host tensors with keys src_img / src_lbl / src_bbox (lists) and
        tar_lbl / tar_bbox.  Yields the host result tensor of every batch, in order."""
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
                    host, fin = self._pending.pop(0)
                    fin.synchronize()
                    yield host
            while self._pending:
                host, fin = self._pending.pop(0)
                fin.synchronize()
                yield host
