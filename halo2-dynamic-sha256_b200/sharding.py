"""Instance sharding across ranks: independent instances, contiguous ranges per rank, no data-path
collective.  After the hot path the ranks all_gather digests + checksums (64 B per instance)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_instances: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [first, last) of rank `rank` out of `world`: sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("bad rank")
    base, rem = divmod(n_instances, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def gather_results(digests, checksums, world: int):
    """all_gather of per-rank digests [n_r*D,32] u8 and checksums [n_r,4] i64 tensors (equal n_r on every rank, as in
    the weak-scaling bench) -> concatenated tensors in rank order + the job checksum (sum of all totals mod 2^64)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        allc, alld = [checksums], [digests]
    else:
        allc = [torch.empty_like(checksums) for _ in range(world)]
        alld = [torch.empty_like(digests) for _ in range(world)]
        dist.all_gather(allc, checksums)
        dist.all_gather(alld, digests)
    cks = torch.cat(allc, 0)
    job = 0
    for v in cks[:, 3].cpu().tolist():
        job = (job + (v & ((1 << 64) - 1))) & ((1 << 64) - 1)
    return torch.cat(alld, 0), cks, job


_nccl = None


def _nccl_lib():
    """libnccl as a ctypes handle: the copy torch has already loaded (same soname), so h2sha_gather's dlopen finds the same one."""
    global _nccl
    if _nccl is None:
        import ctypes as C
        _nccl = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)
    return _nccl


def create_nccl_comm(rank: int, world: int):
    """An `ncclComm_t` of our own (ncclGetUniqueId on rank 0, the 128-byte id shipped through torch.distributed's store,
    ncclCommInitRank on every rank) -- what a Rust / C++ host would create and hand to h2sha_gather."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    class _UniqueId(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]

    L = _nccl_lib()
    L.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
    L.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    uid = _UniqueId()
    if rank == 0:
        rc = L.ncclGetUniqueId(C.byref(uid))
        if rc:
            raise RuntimeError(f"ncclGetUniqueId failed ({rc})")
    box = [C.string_at(C.addressof(uid), 128)] if rank == 0 else [None]
    dist.broadcast_object_list(box, src=0)
    C.memmove(C.addressof(uid), box[0], 128)
    comm = C.c_void_p()
    rc = L.ncclCommInitRank(C.byref(comm), world, uid, rank)
    if rc:
        raise RuntimeError(f"ncclCommInitRank failed ({rc})")
    return comm


def destroy_nccl_comm(comm):
    import ctypes as C
    L = _nccl_lib()
    L.ncclCommDestroy.argtypes = [C.c_void_p]
    L.ncclCommDestroy(comm)


def gather_results_cabi(pkg, digests, checksums, rank: int, world: int, stream):
    """The same gather through the C-ABI: h2sha_gather(ncclComm_t, ...) on `stream` (a torch.cuda.Stream); returns
    (all digests, all checksums, job checksum) like gather_results."""
    import ctypes as C

    import torch
    n = checksums.shape[0]
    D = digests.shape[0] // n
    all_d = torch.empty((world * digests.shape[0], 32), dtype=torch.uint8, device=digests.device)
    all_c = torch.empty((world * n, 4), dtype=torch.int64, device=digests.device)
    comm = create_nccl_comm(rank, world)
    try:
        L = pkg.load_library()
        rc = L.h2sha_gather(comm, n, D, C.c_void_p(digests.data_ptr()), C.c_void_p(checksums.data_ptr()), C.c_void_p(all_d.data_ptr()),
                            C.c_void_p(all_c.data_ptr()), C.c_void_p(stream.cuda_stream))
        if rc:
            raise RuntimeError("h2sha_gather: " + L.h2sha_last_error().decode())
        stream.synchronize()
    finally:
        destroy_nccl_comm(comm)
    job = 0
    for v in all_c[:, 3].cpu().tolist():
        job = (job + (v & ((1 << 64) - 1))) & ((1 << 64) - 1)
    return all_d, all_c, job
