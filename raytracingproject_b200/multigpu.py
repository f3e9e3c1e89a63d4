"""Sample-split multi-GPU rendering: every rank renders the whole
frame for a disjoint range of sample indices - the Sobol index IS the sample number
(kernel/kernel_random.h:75), so the ranges jointly equal the single-device sample
set - and the per-rank films are summed.  Replaces the reference's host-side tile
fan-out (device/device_multi.cpp:374-393, 689-737) and reuses what its "resumable
chunks" do offline (blender/blender_session.cpp:1062-1111).

One process per GPU; the film sum is one NCCL all-reduce over NVLink on the device
film buffer (gloo on CPU tensors in the tests)."""
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


def reduce_film(film, dst=None):
    """Sum the per-rank films in place (all ranks get the sum, or only `dst`)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return film
    if dst is None:
        dist.all_reduce(film, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def display_scale(total_samples):
    """Film sums are normalised only at read-out (kernel_film.h:103-106)."""
    return 1.0 / float(total_samples)
