"""Sharding of independent subcircuits over ranks and the one exchange step of the path.

Mirrors the role of `work()` in /root/reference/mpi-snark/src/bin/node.rs:461-621: subcircuit
proofs are independent within a stage, rank r proves the contiguous chunk
[r * n/W, (r+1) * n/W) (node.rs:471-493 requires n % W == 0), and the only communication is
gathering the fixed-size responses (node.rs:687-715 gather_varcount; here an all_gather of the
proof records over NCCL -- or gloo on CPU in the tests).  No data-path collective exists: a single
MSM or NTT is never split across GPUs.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(num_subcircuits, world_size, rank):
    """Contiguous chunk of subcircuit indices for `rank` (node.rs:471-472 semantics)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    if num_subcircuits % world_size != 0:
        raise ValueError("num_subcircuits must be divisible by the number of workers (node.rs:472)")
    per = num_subcircuits // world_size
    return range(rank * per, (rank + 1) * per)


def gather_records(local_records, world_size=None):
    """all_gather fixed-size per-subcircuit records (one row per local subcircuit) so that every rank
    (in particular the aggregating rank 0) holds all of them in subcircuit order.
    local_records: (n_local, record_words) integer tensor on the rank's device."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if world_size == 1:
        return local_records
    outs = [torch.empty_like(local_records) for _ in range(world_size)]
    dist.all_gather(outs, local_records.contiguous())
    return torch.cat(outs, dim=0)


def derive_stage1_requests(all_commitments):
    """Stub of CoordinatorState::stage_1 (distributed-prover/src/coordinator.rs:315-351): the real coordinator
    builds an inner-product commitment over ALL stage-0 commitments and derives the Fiat-Shamir challenges every
    stage-1 request carries.  The hot path only needs the DEPENDENCY, so the stub folds the gathered records on the
    host into a 64-bit value -- which forces the read-back of every commitment before any proof is enqueued."""
    a = np.ascontiguousarray(all_commitments).view(np.uint8).ravel()
    a = a[: a.size // 8 * 8].view(np.uint64)
    return int(np.bitwise_xor.reduce(a)) if a.size else 0


def stage_barrier(local_commitment_records, world_size=None):
    """The exchange between the two phases of a run (mpi-snark/src/bin/node.rs:488-539: gather stage-0 responses ->
    coordinator.stage_1(&responses) -> scatter stage-1 requests): every rank's commitment records (integer tensor on
    the rank's device) are gathered, brought to the host, and folded into the stage-1 challenge.  Returns
    (challenge, gathered records as a numpy array); identical on every rank."""
    rec = gather_records(local_commitment_records.reshape(1, -1), world_size)
    host = rec.cpu().numpy()                     # synchronises the device stream the gather ran on
    return derive_stage1_requests(host), host
