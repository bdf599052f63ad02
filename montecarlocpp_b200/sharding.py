"""Multi-GPU sharding of FieldProblem::solve: phonons are independent histories (the reference already
splits them over OpenMP threads with a static contiguous partition, problem.cpp:383-384, and combines the
per-thread fields with one sum, main.cpp:162-165).  Rank g of G owns particle ids
[g*ceil(N/G), min(N, (g+1)*ceil(N/G))); the Philox key is the GLOBAL particle id, so the result does not
depend on G (up to fp summation order).  The only data-path collective is one all-reduce (sum) of the raw
rows x cols fp64 tally per solve; normalisation (problem.cpp:439-444) is applied after it."""


def shard_range(nemit, world, rank):
    per = (nemit + world - 1) // world
    return min(nemit, rank * per), min(nemit, (rank + 1) * per)


def allreduce_raw_field(raw):
    """Sum the raw tally over ranks in place (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(raw, op=dist.ReduceOp.SUM)
    return raw
