"""Sample-split multi-GPU rendering: every rank renders the whole
frame for a disjoint range of sample indices - the Sobol index IS the sample number
(kernel/kernel_random.h:75), so the ranges jointly equal the single-device sample
set - and the per-rank films are summed.  Replaces the reference's host-side tile
fan-out (device/device_multi.cpp:374-393, 689-737) and reuses what its "resumable
chunks" do offline (blender/blender_session.cpp:1062-1111).

One process per GPU; the film sum is one NCCL all-reduce over NVLink on the device
film buffer (gloo on CPU tensors in the tests).  `FilmReducer` takes the sum off the
render stream: frames alternate between two film buffers and the all-reduce of frame k
runs on a side stream while frame k+1 is being traced, so a rank that finished early
starts its next frame instead of waiting for the slowest one at every frame."""
import torch
import torch.distributed as dist


def weak_range(rank, spp_per_rank, start_sample=0):
    """Fixed work per rank: rank r renders [start + r*spp, start + (r+1)*spp)."""
    return start_sample + rank * spp_per_rank, spp_per_rank


def strong_range(rank, world, total_spp, start_sample=0):
    """Fixed total work: [start, start+total) cut into `world` contiguous ranges whose
    sizes differ by at most one sample."""
    base, rem = divmod(total_spp, world)
    begin = start_sample + rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def _distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def reduce_film(film, dst=None):
    """Sum the per-rank films in place (all ranks get the sum, or only `dst`)."""
    if not _distributed():
        return film
    if dst is None:
        dist.all_reduce(film, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def display_scale(total_samples):
    """Film sums are normalised only at read-out (kernel_film.h:103-106)."""
    return 1.0 / float(total_samples)


class FilmReducer:
    """Double-buffered film sum.

        film = reducer.begin_frame()        # zeroed buffer, safe to render into
        ... render into film on the current stream ...
        reducer.end_frame()                 # all-reduce on the side stream
        ...
        summed = reducer.finish()           # last frame's sum, current stream waits for it

    On CUDA tensors the reduce runs on its own stream, ordered after the frame's render by
    an event, and is timed by its own pair of events (`reduce_ms()`): that time is the
    collective plus the wait for the slowest rank of that frame.  On CPU tensors (gloo, the
    tests) it degrades to a plain in-place all-reduce with the same call sequence."""

    def __init__(self, numel, device, dtype=torch.float32, timed=True):
        self.films = [torch.zeros(numel, dtype=dtype, device=device) for _ in range(2)]
        self.cuda = self.films[0].is_cuda
        self.frame = 0
        self.cur = None
        self.timed = timed and self.cuda
        self._pairs = []
        if self.cuda:
            self.side = torch.cuda.Stream(device=device)
            self.rendered = [torch.cuda.Event() for _ in range(2)]
            self.reduced = [None, None]

    def begin_frame(self):
        b = self.frame & 1
        self.cur = self.films[b]
        if self.cuda and self.reduced[b] is not None:
            torch.cuda.current_stream().wait_event(self.reduced[b])
        self.cur.zero_()
        return self.cur

    def end_frame(self):
        b = self.frame & 1
        self.frame += 1
        if not self.cuda:
            reduce_film(self.cur)
            return
        self.rendered[b].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.rendered[b])
            if self.timed:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(self.side)
            reduce_film(self.cur)
            if self.timed:
                e1.record(self.side)
                self._pairs.append((e0, e1))
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.reduced[b] = ev

    def finish(self):
        """Makes the current stream wait for every outstanding reduce; returns the film of
        the last frame."""
        if self.cuda:
            for ev in self.reduced:
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
        return self.cur

    def reduce_ms(self):
        """Average device time of the timed reduces since the last call (after a
        synchronize)."""
        if not self._pairs:
            return 0.0
        ms = sum(a.elapsed_time(b) for a, b in self._pairs) / len(self._pairs)
        self._pairs = []
        return ms
