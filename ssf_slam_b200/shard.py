"""Multi-GPU plumbing: independent sequences shard across ranks (one process per GPU); the only communication is
a gather of per-frame poses and masks (the reference itself has no collective).

The gather is ONE collective per call (poses and masks packed into one byte buffer -> ``all_gather_into_tensor``: NCCL over
NVLink on GPUs, gloo on CPU) and is meant to run off the critical path: ``ResultGatherer`` issues it on its own CUDA stream
behind an event of the producing stream, so the next batch's kernels never wait for it."""
import torch
import torch.distributed as dist


def my_sequences(n_sequences, rank, world):
    """Sequence ids owned by ``rank``: seq_id % world == rank (config 4: 64 sequences over 1/2/4/8 GPUs)."""
    return [s for s in range(n_sequences) if s % world == rank]


def _pack(odom, mask):
    """odom f64 [..., 7], mask u8 [..., N] -> one contiguous u8 buffer (odom bytes first)."""
    return torch.cat([odom.contiguous().view(torch.uint8).reshape(-1), mask.contiguous().reshape(-1)])


def _unpack(buf, world, odom_shape, mask_shape):
    n_od = 8
    for d in odom_shape:
        n_od *= d
    buf = buf.view(world, -1)
    od = buf[:, :n_od].contiguous().view(torch.float64).view((world,) + tuple(odom_shape))
    mk = buf[:, n_od:].contiguous().view((world,) + tuple(mask_shape))
    return od, mk


def gather_results(odom, mask, group=None):
    """All ranks contribute odom f64 [..., 7] and mask u8 [..., N]; every rank returns the rank-major concatenation
    ([world, ...] each), through one ``all_gather_into_tensor`` of the packed bytes.  ~8 KB per frame: latency-bound."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return odom.unsqueeze(0), mask.unsqueeze(0)
    world = dist.get_world_size(group)
    packed = _pack(odom, mask)
    out = torch.empty(world * packed.numel(), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return _unpack(out, world, odom.shape, mask.shape)


class ResultGatherer:
    """Gathers each finished batch's poses and masks on a private CUDA stream: ``push(odom, mask, producer_stream)`` records
    an event on the producing stream and queues the collective behind it, so it overlaps the following batches; ``finish()``
    makes the caller's stream wait for every queued gather and returns them.  CUDA events around every collective give the
    time the communication itself took (``gather_ms``), which is reported next to, not inside, the kernel time."""

    def __init__(self, device, group=None):
        self.device, self.group = device, group
        self.stream = torch.cuda.Stream(device=device)
        self._out, self._ev = [], []

    def push(self, odom, mask, producer_stream):
        done = torch.cuda.Event()
        done.record(producer_stream)
        self.stream.wait_event(done)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            odom.record_stream(self.stream)
            mask.record_stream(self.stream)
            e0.record(self.stream)
            self._out.append(gather_results(odom, mask, self.group))
            e1.record(self.stream)
        self._ev.append((e0, e1))

    def finish(self, consumer_stream=None):
        (consumer_stream or torch.cuda.current_stream(self.device)).wait_stream(self.stream)
        out, self._out = self._out, []
        return out

    def gather_ms(self):
        """Sum of the collectives' durations (call after a device synchronize); clears the event list."""
        ms = sum(a.elapsed_time(b) for a, b in self._ev)
        self._ev = []
        return ms
