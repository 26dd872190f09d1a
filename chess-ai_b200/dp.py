"""Data-parallel plumbing: one process per GPU, torch.distributed only for rendezvous (broadcast of the NCCL unique
id); the gradient all-reduce itself runs inside libchessai_b200.so on the library's own NCCL communicator.

Evaluation needs none of this (boards shard with no collective). Training: every rank calls
DataParallel.train_batch on ITS shard; the library all-reduces (sum dW, sum E, count) and then every rank applies the
identical update and takes the identical accept / reject decision.
"""
import ctypes as C

import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced, disjoint cover of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object of known length from `src` with whatever torch.distributed backend is up."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


class DataParallel:
    """Wraps a Network for data-parallel minibatch training on the current torch.distributed world."""

    def __init__(self, cab, network, peer=False):
        """peer=False: gradient all-reduce by NCCL inside the library. peer=True: the ranks' gradient buffers are mapped
        into each other (CUDA IPC over NVLink peer access) and the reduction is fused into the update kernel."""
        import torch.distributed as dist
        self.cab, self.net, self.peer = cab, network, peer
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        lib = cab.lib()
        if peer:
            mine = (C.c_ubyte * 64)()
            cab._ck(lib.cab_dp_peer_export(network._h, mine))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine))
            buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
            err = None
            try:
                cab._ck(lib.cab_dp_peer_attach(network._h, buf, self.rank, self.world))
            except Exception as exc:            # e.g. no peer access between two of the devices
                err = str(exc)
            errs = [None] * self.world          # all ranks agree on the outcome (this is also the barrier: nobody steps
            dist.all_gather_object(errs, err)   # before every rank has mapped every block)
            if any(e is not None for e in errs):
                lib.cab_dp_shutdown(network._h)
                raise RuntimeError("peer-memory data parallel unavailable: %s" % next(e for e in errs if e is not None))
            return
        ident = (C.c_ubyte * 128)()
        if self.rank == 0:
            cab._ck(lib.cab_dp_unique_id(ident))
        ident_bytes = broadcast_bytes(bytes(ident), src=0)
        buf = (C.c_ubyte * 128).from_buffer_copy(ident_bytes)
        cab._ck(lib.cab_dp_init(network._h, buf, self.rank, self.world))

    def train_batch(self, codes, targets, lam, max_const=15.0, rate=1.1):
        """This rank's shard (host arrays). Returns (lambda, accepted, E, E') — identical on every rank."""
        return self.cab.Optimizer.train_batch(self.net, codes, targets, lam, max_const, rate)

    def train_batch_dev(self, codes_ptr, targets_ptr, n, lam, max_const=15.0, rate=1.1, stream=0):
        lam_c, acc, e0, e1 = C.c_double(lam), C.c_int(0), C.c_double(0), C.c_double(0)
        self.cab._ck(self.cab.lib().cab_train_batch_dev(self.net._h, codes_ptr, targets_ptr, n, C.byref(lam_c), max_const, rate,
                                                         C.byref(acc), C.byref(e0), C.byref(e1), stream))
        return lam_c.value, bool(acc.value), e0.value, e1.value

    def shutdown(self):
        if self.peer:
            import torch
            import torch.distributed as dist
            torch.cuda.synchronize()
            dist.barrier()                      # nobody unmaps a block a peer may still read
        self.cab._ck(self.cab.lib().cab_dp_shutdown(self.net._h))
